#!/usr/bin/env python3
"""Walk a zstd frame and print its blocks: type, sizes, literals section, number of sequences, table modes (RFC 8878).  Test tool."""
import sys


def walk(frame, verbose=True):
    p = 0
    assert frame[:4] == b"\x28\xb5\x2f\xfd"
    fhd = frame[4]
    single = (fhd >> 5) & 1
    p = 5 + (0 if single else 1) + (0, 1, 2, 4)[fhd & 3]
    nb = (1 if single else 0, 2, 4, 8)[fhd >> 6]
    fcs = int.from_bytes(frame[p:p + nb], "little") + (256 if nb == 2 else 0)
    p += nb
    out = []
    while True:
        bh = int.from_bytes(frame[p:p + 3], "little")
        last, typ, bs = bh & 1, (bh >> 1) & 3, bh >> 3
        p += 3
        info = {"type": ("raw", "rle", "cmp")[typ], "size": bs}
        if typ == 2:
            b0 = frame[p]
            lt, sf = b0 & 3, (b0 >> 2) & 3
            if lt < 2:
                if sf in (0, 2):
                    hdr, regen = 1, b0 >> 3
                elif sf == 1:
                    hdr, regen = 2, int.from_bytes(frame[p:p + 2], "little") >> 4
                else:
                    hdr, regen = 3, int.from_bytes(frame[p:p + 3], "little") >> 4
                sec = hdr + (regen if lt == 0 else 1)
            else:
                if sf <= 1:
                    v = int.from_bytes(frame[p:p + 3], "little"); hdr, regen, comp = 3, (v >> 4) & 0x3FF, (v >> 14) & 0x3FF
                elif sf == 2:
                    v = int.from_bytes(frame[p:p + 4], "little"); hdr, regen, comp = 4, (v >> 4) & 0x3FFF, (v >> 18) & 0x3FFF
                else:
                    v = int.from_bytes(frame[p:p + 5], "little"); hdr, regen, comp = 5, (v >> 4) & 0x3FFFF, (v >> 22) & 0x3FFFF
                sec = hdr + comp
            q = p + sec
            n0 = frame[q]
            if n0 == 0:
                nseq, h = 0, 1
            elif n0 < 128:
                nseq, h = n0, 1
            elif n0 < 255:
                nseq, h = ((n0 - 128) << 8) + frame[q + 1], 2
            else:
                nseq, h = frame[q + 1] + (frame[q + 2] << 8) + 0x7F00, 3
            info.update({"lit": ("raw", "rle", "huf", "treeless")[lt], "lit_regen": regen, "lit_sec": sec, "nseq": nseq, "seq_sec": bs - sec,
                         "modes": ("%02x" % frame[q + h]) if nseq else ""})
            p += bs
        elif typ == 1:
            p += 1
        else:
            p += bs
        out.append(info)
        if verbose:
            print(info)
        if last:
            break
    return fcs, out


if __name__ == "__main__":
    walk(open(sys.argv[1], "rb").read())
