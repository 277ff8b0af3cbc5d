#!/usr/bin/env python3
"""Regenerates the committed golden vectors from the reference's own fixtures.

Run in the build container only (needs /root/reference, read-only):

    python tests/golden/make_golden.py

What it extracts, without libhdf5 (absent from the image):
  * every chunk payload of  test_data/multi_fast5_vbz.fast5     (filter 32020, cd_values 0,2,1,1,1)
  * every chunk payload of  test_data/multi_fast5_vbz_v1.fast5  (cd_values 1,2,1,1)
  * the matching zlib chunks of test_data/multi_fast5_zip.fast5 -> expected int16 samples
    (restates python/test/test_vbz_filter.py:57-73: Raw/Signal must be equal across the three files)

HDF5 layout facts used (superblock v0 files, 1-D chunked datasets, one chunk per read):
  v1 B-tree leaf: "TREE", u8 type(=1 chunk), u8 level(=0), u16 entries, u64 left, u64 right, then per
  entry key {u32 chunk_bytes, u32 filter_mask, u64 offset[2]} followed by u64 child (chunk) address.

Outputs (committed):
  tests/golden/fast5_chunks.npz   - vbz payloads (v0 file), flags saying the v1 file's payloads are
                                    byte-identical, zlib payloads, per-read sha256 of samples
  tests/golden/fuzz_corpus.npz    - the reference's libFuzzer corpus (vbz/fuzzing/fuzz_corpus), packed
  tests/golden/kat.json           - byte-exact known-answer vectors restated from the reference tests
                                    (vbz/test/vbz_test.cpp:204,238,321; vbz/test/streamvbyte_test.cpp:207-236)
                                    plus extra probes whose expected bytes are the compiled
                                    reference's output at generation time.
"""
import ctypes, hashlib, json, os, struct, sys, zlib
import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def chunks_of(path):
    blob = open(path, "rb").read()
    out = []
    pos = 0
    while True:
        pos = blob.find(b"TREE", pos)
        if pos < 0:
            break
        ntype, level, used = struct.unpack_from("<BBH", blob, pos + 4)
        if ntype == 1 and level == 0 and 0 < used < 64:
            p = pos + 24
            ok = True
            ents = []
            for _ in range(used):
                size, mask, off0, off1, addr = struct.unpack_from("<IIQQQ", blob, p)
                p += 32
                if addr + size > len(blob) or size == 0:
                    ok = False
                    break
                ents.append((off0, blob[addr:addr + size]))
            if ok:
                out.extend(e[1] for e in ents)
        pos += 4
    return out


def filter_cd_values(path):
    blob = open(path, "rb").read()
    i = blob.find(b"vbz\0")
    if i < 0:
        return None
    fid, namelen, flags, ncd = struct.unpack_from("<HHHH", blob, i - 8)
    assert fid == 32020, fid
    return list(struct.unpack_from("<%dI" % ncd, blob, i + namelen))


class Opts(ctypes.Structure):
    _fields_ = [("zz", ctypes.c_bool), ("isz", ctypes.c_uint), ("lvl", ctypes.c_uint), ("ver", ctypes.c_uint)]


def ref_lib():
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libvbz_ref.so"))
    for n in ("vbz_compress_sized", "vbz_decompress_sized", "vbz_compress", "vbz_decompress"):
        getattr(lib, n).restype = ctypes.c_uint32
        getattr(lib, n).argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_uint32, ctypes.POINTER(Opts)]
    lib.vbz_max_compressed_size.restype = ctypes.c_uint32
    lib.vbz_max_compressed_size.argtypes = [ctypes.c_uint32, ctypes.POINTER(Opts)]
    return lib


def ref_compress(lib, arr, zz, isz, lvl, ver, sized):
    raw = np.ascontiguousarray(arr).view(np.uint8)
    o = Opts(zz, isz, lvl, ver)
    cap = lib.vbz_max_compressed_size(raw.size, ctypes.byref(o))
    dst = np.zeros(cap, np.uint8)
    fn = lib.vbz_compress_sized if sized else lib.vbz_compress
    n = fn(raw.ctypes.data, raw.size, dst.ctypes.data, cap, ctypes.byref(o))
    assert n < 0xFFFFFFF9, n
    return dst[:n].tolist()


def main():
    lib = ref_lib()
    v0 = chunks_of(f"{REF}/test_data/multi_fast5_vbz.fast5")
    v1 = chunks_of(f"{REF}/test_data/multi_fast5_vbz_v1.fast5")
    zp = chunks_of(f"{REF}/test_data/multi_fast5_zip.fast5")
    # Raw/Signal chunks are the big ones; other chunked datasets (if any) are tiny.
    zsig = {}
    for c in zp:
        try:
            raw = zlib.decompress(c)
        except zlib.error:
            continue
        if len(raw) >= 2 * 5000:
            zsig[len(raw)] = (c, raw)
    def sig_chunks(chs):
        sel = {}
        for c in chs:
            if len(c) < 13:
                continue
            n = struct.unpack_from("<I", c, 0)[0]
            if n in zsig and c[4:8] == b"\x28\xb5\x2f\xfd":
                sel[n] = c
        return sel
    s0, s1 = sig_chunks(v0), sig_chunks(v1)
    assert len(s0) == 10 and len(s1) == 10 and set(s0) == set(s1) == set(zsig), (len(s0), len(s1), len(zsig))
    sizes = sorted(s0)
    # cross-check with the compiled reference before committing anything
    for n in sizes:
        for ver, c in ((0, s0[n]), (1, s1[n])):
            o = Opts(True, 2, 1, ver)
            dst = np.zeros(n, np.uint8)
            src = np.frombuffer(c, np.uint8)
            r = lib.vbz_decompress_sized(src.ctypes.data, src.size, dst.ctypes.data, n, ctypes.byref(o))
            assert r == n and dst.tobytes() == zsig[n][1], (n, ver, r)
    cd0 = filter_cd_values(f"{REF}/test_data/multi_fast5_vbz.fast5")
    cd1 = filter_cd_values(f"{REF}/test_data/multi_fast5_vbz_v1.fast5")
    np.savez(
        os.path.join(HERE, "fast5_chunks.npz"),
        raw_bytes=np.array(sizes, np.int64),
        vbz_payload=np.frombuffer(b"".join(s0[n] for n in sizes), np.uint8),
        vbz_payload_len=np.array([len(s0[n]) for n in sizes], np.int64),
        v1_identical=np.array([s0[n] == s1[n] for n in sizes]),
        zip_payload=np.frombuffer(b"".join(zsig[n][0] for n in sizes), np.uint8),
        zip_payload_len=np.array([len(zsig[n][0]) for n in sizes], np.int64),
        sample_sha256=np.array([hashlib.sha256(zsig[n][1]).hexdigest() for n in sizes]),
        cd_values_v0=np.array(cd0, np.int64), cd_values_v1=np.array(cd1, np.int64),
    )
    print("reads:", [n // 2 for n in sizes], "vbz bytes", sum(len(s0[n]) for n in sizes), "v1 identical", all(s0[n] == s1[n] for n in sizes), cd0, cd1)

    # ---- known-answer vectors ------------------------------------------------
    kat = []
    def add(name, dtype, values, zz, isz, lvl, ver, sized, expect=None, source=""):
        arr = np.array(values, dtype=dtype)
        got = ref_compress(lib, arr, zz, isz, lvl, ver, sized)
        if expect is not None:
            e = [x & 0xFF for x in expect]
            assert got == e, (name, got, e)
        kat.append(dict(name=name, dtype=dtype, values=[int(v) for v in arr.tolist()], zz=zz, isz=isz, lvl=lvl, ver=ver,
                        sized=sized, bytes=got, source=source or "compiled reference output (oracle/_ref, x86 -mssse3, libzstd 1.5.5)"))
    add("int32_zz_lvl0", "int32", [5, 4, 3, 2, 1], True, 4, 0, 0, False, [0, 0, 10, 1, 1, 1, 1], "vbz/test/vbz_test.cpp:204")
    add("int32_zz_lvl100", "int32", [5, 4, 3, 2, 1], True, 4, 100, 0, False,
        [40, -75, 47, -3, 32, 7, 57, 0, 0, 0, 0, 10, 1, 1, 1, 1], "vbz/test/vbz_test.cpp:238")
    add("int32_zz_lvl0_sized", "int32", [5, 4, 3, 2, 1], True, 4, 0, 0, True, [20, 0, 0, 0, 0, 0, 10, 1, 1, 1, 1], "vbz/test/vbz_test.cpp:321")
    sq = [0, -1, 4, -9, 16, -25, 36, -49, 64, -81, 100]
    for ver in (0, 1):
        add(f"int16_zz_squares_v{ver}", "int16", sq, True, 2, 0, ver, False,
            [0, 0, 20, 0, 1, 10, 25, 50, 81, 122, -87, -30, 33, 1, 106, 1], "vbz/test/streamvbyte_test.cpp:207-219")
        add(f"uint16_nozz_squares_v{ver}", "uint16", [i * i for i in range(11)], False, 2, 0, ver, False,
            [0, 0, 0] + [i * i for i in range(11)], "vbz/test/streamvbyte_test.cpp:225-236")
    add("int16_zz_fullrange", "int16", [32767, -32768, 32767, -32768, 0, 0, 0, 0, 32767, -32768], True, 2, 0, 0, False)
    add("int16_nozz_neg", "int16", [-1, 1, -2, 2], False, 2, 0, 0, False)
    add("int8_zz_v1", "int8", [0, 1, 2, 3, 20, -20, 100, -100, 5], True, 1, 0, 1, False)
    add("int8_zz_v0", "int8", [0, 1, 2, 3, 20, -20, 100, -100, 5], True, 1, 0, 0, False)
    add("int8_nozz_v1", "int8", [0, 1, 2, 3, 20, -20, 100, -100, 5], False, 1, 0, 1, False)
    add("int8_nozz_v0", "int8", [0, 1, 2, 3, 20, -20, 100, -100, 5], False, 1, 0, 0, False)
    add("int32_zz_extremes", "int32", [7, -2**31, 2**31 - 1, 0, 1], True, 4, 0, 0, False)
    add("int32_nozz_mixed", "int32", [0, 255, 256, 65535, 65536, 2**24 - 1, 2**24, -1, 3], False, 4, 0, 0, False)
    add("empty_lvl0_sized", "int16", [], True, 2, 0, 0, True, [0, 0, 0, 0])
    add("empty_lvl1_sized", "int16", [], True, 2, 1, 1, True, [0, 0, 0, 0, 40, 181, 47, 253, 32, 0, 1, 0, 0])
    for n in (1, 2, 3, 4, 5, 7, 8, 9, 15, 16, 17, 31, 33):
        add(f"int16_zz_ramp{n}", "int16", [(-1) ** i * i * 37 for i in range(n)], True, 2, 0, 0, True)
    json.dump(kat, open(os.path.join(HERE, "kat.json"), "w"), indent=0)
    print(len(kat), "KATs written")

    # ---- the reference's fuzz corpus (vbz/fuzzing/fuzz_corpus/*, 238 files, 60 KB) -----------------
    cdir = f"{REF}/vbz/fuzzing/fuzz_corpus"
    names = sorted(os.listdir(cdir))
    blobs = [open(os.path.join(cdir, n), "rb").read() for n in names]
    np.savez_compressed(os.path.join(HERE, "fuzz_corpus.npz"),
                        data=np.frombuffer(b"".join(blobs), np.uint8),
                        lens=np.array([len(b) for b in blobs], np.int64), names=np.array(names))
    print(len(names), "fuzz corpus files,", sum(map(len, blobs)), "bytes")


if __name__ == "__main__":
    main()
