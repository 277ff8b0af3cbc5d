"""Stage 2, decode direction: the product's zstd frame decoder (csrc/zstd_dec_kernels.cuh) against frames written by
the reference (fast5 fixtures) and by libzstd at many levels, through the vbz C API (zstd_compression_level != 0).
Runs on the CUDA emulator build (CPU suite) and on the real library (-m gpu)."""
import numpy as np
import pytest

import data_gen
import golden_data
import oracle_lib as ol
import product
from oracle_lib import Options

port = ol.port()


@pytest.fixture(params=product.BACKENDS)
def backend(request):
    return product.get(request.param)


def test_reference_written_chunks_decode(backend):
    """The reference's own compressed files must decode bit-exactly: v0 and v1 fixtures."""
    capi = backend[0]
    reads = golden_data.fast5_reads()
    take = reads if not capi.path.endswith("cpusim.so") else reads[:4]
    for rd in take:
        want = rd["samples"]
        assert rd["v1_identical"]  # multi_fast5_vbz_v1.fast5 stores byte-identical chunk payloads (make_golden.py checks it)
        for ver in (0, 1):
            r, got = capi.decompress(rd["vbz_chunk"], Options(True, 2, 1, ver), want.nbytes, sized=True)
            assert r == want.nbytes, (ver, r)
            assert got == want.tobytes(), ver


def _payloads(rng):
    yield b""
    yield b"a"
    yield bytes(1000)
    yield bytes(rng.integers(0, 256, 5000, dtype=np.uint8))                  # incompressible -> raw blocks
    yield bytes(rng.integers(0, 4, 40000, dtype=np.uint8))                   # small alphabet -> direct Huffman weights
    yield bytes(np.repeat(rng.integers(0, 256, 300, dtype=np.uint8), 50))    # runs -> matches with small offsets
    yield (b"the quick brown fox jumps over the lazy dog. " * 900)           # long-range repeats, repeat offsets
    yield ol.stage1_encode(data_gen.nanopore_like(60000, rng), 2, True)      # what VBZ feeds zstd
    yield bytes(np.arange(200000, dtype=np.uint32).view(np.uint8))           # several blocks, structured
    yield bytes(rng.integers(0, 256, 140000, dtype=np.uint8)) + bytes(140000)  # raw block then RLE-ish block


@pytest.mark.parametrize("level", [1, 3, 9, 19])
def test_libzstd_frames_decode(backend, level):
    capi = backend[0]
    rng = np.random.default_rng(level)
    o = Options(False, 0, level, 0)  # zstd only: vbz_decompress == ZSTD_decompress into the destination
    for data in _payloads(rng):
        if capi.path.endswith("cpusim.so") and len(data) > 150000 and level > 3:
            continue
        frame = ol.zstd_compress_lib(data, level)
        r, got = capi.decompress(frame, o, len(data), sized=False)
        assert r == len(data), (level, len(data), r)
        assert got == data
        # capacity smaller than the content size (vbz.cpp:256-259) and larger
        if len(data) > 1:
            assert capi.decompress(frame, o, len(data) - 1, sized=False)[0] == ol.ERR_DEST_SIZE
        r2, got2 = capi.decompress(frame, o, len(data) + 7, sized=False)
        assert r2 == len(data) and got2 == data


def test_full_vbz_decode_matches_oracle(backend):
    capi = backend[0]
    rng = np.random.default_rng(9)
    for dtype, o in ((np.int16, Options(True, 2, 1, 0)), (np.int16, Options(True, 2, 3, 1)), (np.int32, Options(True, 4, 1, 0)),
                     (np.uint32, Options(False, 4, 3, 0)), (np.int8, Options(True, 1, 1, 0)), (np.int8, Options(True, 1, 1, 1)), (np.uint8, Options(False, 1, 1, 1))):
        for n in (0, 1, 100, 5000, 30000):
            arr = data_gen.nanopore_like(n, rng).astype(dtype) if dtype != np.uint32 else rng.integers(0, 1 << 20, n).astype(np.uint32)
            for sized in (True, False):
                rc, blob = port.compress(arr, o, sized=sized)
                assert not ol.is_error(rc)
                a = capi.decompress(blob, o, arr.nbytes, sized=sized)
                b = port.decompress(blob, o, arr.nbytes, sized=sized)
                assert a == b and a[1] == arr.tobytes(), (dtype, n, sized, a[0], b[0])


def test_damaged_frames_report_like_the_oracle(backend):
    capi = backend[0]
    rng = np.random.default_rng(4)
    arr = data_gen.nanopore_like(20000, rng)
    o = Options(True, 2, 1, 0)
    _, good = port.compress(arr, o, sized=True)
    cases = [good[:10], good[:-1], good[:len(good) // 2], good + b"\\0", good[:4] + b"\\0\\0\\0\\0" + good[8:]]
    for pos in (4, 5, 8, 9, 12, 20, 200, len(good) - 2):
        bad = bytearray(good)
        bad[pos] ^= 0x55
        cases.append(bytes(bad))
    for blob in cases:
        for oo, ds in ((o, arr.nbytes), (Options(True, 2, 1, 2), arr.nbytes), (o, arr.nbytes - 1), (Options(False, 0, 1, 0), arr.nbytes)):
            a = capi.decompress(blob, oo, ds, sized=True)
            b = port.decompress(blob, oo, ds, sized=True)
            assert a[0] == b[0], (len(blob), oo.vbz_version, oo.integer_size, ds, a[0], b[0])
            if not ol.is_error(b[0]):
                assert a[1] == b[1]


def test_bit_flips_in_own_frames_never_decode_to_something_else(backend):
    """Single-bit damage anywhere in a frame of this library's own writer (S groups: sequences sections, FSE tables, Repeat_Mode blocks,
    treeless literals): the block-parallel decoder (sets of blocks a lane each, literals expanded in place, chains handed from warp to
    warp) must end - with an error, or with exactly what the oracle makes of the same bytes.  Never more lenient than the oracle."""
    capi, batch, _ = backend
    rng = np.random.default_rng(9)
    small = capi.path.endswith("cpusim.so")
    arr = np.concatenate([data_gen.nanopore_like(9000, rng), np.full(3000, 417, np.int16), data_gen.nanopore_like(4000, rng)])
    o = Options(True, 2, 1, 1)
    (code, good), = batch.compress([arr], o, sized=True)
    assert code == len(good)
    nbits = 8 * len(good)
    flips = list(range(32, nbits, 97 if small else 13))
    blobs = []
    for bit in flips:
        bad = bytearray(good)
        bad[bit >> 3] ^= 1 << (bit & 7)
        blobs.append(bytes(bad))
    res = batch.decompress(blobs, [arr.nbytes] * len(blobs), o, sized=True)  # one batch: a damaged read must not touch its neighbours either
    lenient = 0
    for bit, blob, (rc, out) in zip(flips, blobs, res):
        rp, want = port.decompress(blob, o, arr.nbytes, sized=True)
        if ol.is_error(rc):
            continue
        assert not ol.is_error(rp), (bit, rc, rp)
        assert rc == rp and out == want, (bit, rc, rp)
        lenient += 1
    (rc, out), = batch.decompress([good], [arr.nbytes], o, sized=True)
    assert rc == arr.nbytes and out == arr.tobytes()


def test_fuzz_corpus_level1_decode(backend):
    """Restates the decode half of vbz/fuzzing/vbz_fuzz.cpp:138-161 at zstd level 1: arbitrary bytes must be rejected
    (or decoded) exactly as the oracle does."""
    capi = backend[0]
    corpus = golden_data.fuzz_corpus()
    step = 1 if not capi.path.endswith("cpusim.so") else 5
    for blob in corpus[::step]:
        n = len(blob)
        for o in (Options(True, 2, 1, 0), Options(False, 4, 1, 0), Options(True, 1, 1, 0), Options(True, 1, 1, 1), Options(False, 0, 1, 0)):
            for sized in (True, False):
                for ds in sorted(set([0, 4, 64, n, 4 * n])):
                    a = capi.decompress(blob, o, ds, sized=sized)
                    b = port.decompress(blob, o, ds, sized=sized)
                    assert a[0] == b[0], (n, ds, sized, o.integer_size, a[0], b[0])
                    if not ol.is_error(b[0]):
                        assert a[1] == b[1]


def _decode_stats(backend):
    import ctypes
    _, batch, lib = backend
    out = (ctypes.c_uint32 * 4)()
    assert lib.vbz_b200_decode_stats(batch.ctx, out) == 0
    return {"units": out[0], "general": out[1], "seq_frames": out[2], "records": out[3]}


def test_frames_with_sequences_take_the_block_parallel_path(backend):
    """What real fast5 files contain - frames written by the reference's libzstd, whose first block carries sequences - must not
    fall back to the serial warp-per-frame kernel: the walker lists their blocks, the sequence kernels handle the blocks with
    sequences (csrc/zstd_dec_kernels.cuh)."""
    _, batch, _ = backend
    small = backend[0].path.endswith("cpusim.so")
    reads = golden_data.fast5_reads()
    reads = reads[:5] if small else reads
    o = Options(True, 2, 1, 1)
    res = batch.decompress([rd["vbz_chunk"] for rd in reads], [rd["samples"].nbytes for rd in reads], o, sized=True)
    for rd, (code, blob) in zip(reads, res):
        assert code == rd["samples"].nbytes and blob == rd["samples"].tobytes()
    st = _decode_stats(backend)
    assert st["general"] == 0 and st["seq_frames"] == len(reads) and st["records"] > 0, st
    # libzstd at several levels on inputs with matches of every kind: repeat offsets, long matches, several blocks with sequences
    rng = np.random.default_rng(5)
    datas = [bytes(np.repeat(rng.integers(0, 256, 300, dtype=np.uint8), 50)), b"the quick brown fox jumps over the lazy dog. " * 900,
             bytes(np.arange(60000 if small else 200000, dtype=np.uint32).view(np.uint8)), ol.stage1_encode(data_gen.nanopore_like(60000, rng), 2, True),
             ol.stage1_encode(np.tile(data_gen.nanopore_like(2000, rng), 40 if small else 150), 2, True)]
    for level in (1, 3, 9) if small else (1, 2, 3, 5, 9, 19):
        frames = [ol.zstd_compress_lib(d, level) for d in datas]
        res = batch.decompress(frames, [len(d) for d in datas], Options(False, 0, level, 0), sized=False)
        for d, (code, blob) in zip(datas, res):
            assert code == len(d) and blob == d, (level, len(d), code)
        st = _decode_stats(backend)
        assert st["general"] == 0, (level, st)


def test_own_frames_take_the_block_parallel_path_one_at_a_time(backend):
    """A read decoded on its own (the H5Z filter's case: one chunk per call) has no neighbours whose unit slots it could borrow: the slots
    reserved per frame (csrc/svb_types.h: zstd_units_of) must cover what this library's writer cuts - 2 KiB literal blocks in the S groups
    of the control-byte region included.  A frame that ends up in the general kernel costs 20 ms instead of 0.2."""
    capi, batch, _ = backend
    rng = np.random.default_rng(8)
    small = capi.path.endswith("cpusim.so")
    o = Options(True, 2, 1, 1)
    for n in (140000, 200000) if small else (140000, 300000, 700000, 1000000):
        arr = data_gen.nanopore_like(n, rng)
        (code, blob), = batch.compress([arr], o, sized=True)
        assert code == len(blob)
        (rc, back), = batch.decompress([blob], [arr.nbytes], o, sized=True)
        assert rc == arr.nbytes and back == arr.tobytes()
        st = _decode_stats(backend)
        assert st["general"] == 0 and st["seq_frames"] == 1, (n, st)


def test_pool_overflow_and_damage_fall_back_to_the_general_kernel(backend):
    """A frame the sequence kernel cannot vouch for is decided by the general kernel; its neighbours in the batch are not affected."""
    _, batch, _ = backend
    rng = np.random.default_rng(6)
    good = b"abcdefgh" * 4000
    frame = ol.zstd_compress_lib(good, 3)
    bad = bytearray(frame)
    bad[len(bad) // 2] ^= 0x10
    sig = ol.stage1_encode(data_gen.nanopore_like(30000, rng), 2, True)
    frames = [frame, bytes(bad), ol.zstd_compress_lib(sig, 1), frame[:-3]]
    sizes = [len(good), len(good), len(sig), len(good)]
    res = batch.decompress(frames, sizes, Options(False, 0, 1, 0), sized=False)
    want = [port.decompress(f, Options(False, 0, 1, 0), s, sized=False) for f, s in zip(frames, sizes)]
    for (code, blob), (wc, wb) in zip(res, want):
        assert code == wc and (ol.is_error(wc) or blob == wb)


def test_blocks_that_repeat_tables_ride_with_the_block_before_them(backend):
    """The sequence kernel takes a warp per CHAIN: a block that brings all three tables and the blocks behind it that repeat any of them
    (Repeat_Mode, what libzstd's lazy strategies - level 6 and up - choose block after block); repeat-offset codes cross the chains and
    are settled by whoever finishes the frame's last one.  Frames with every mix of both, several per batch, next to own frames."""
    _, batch, _ = backend
    small = backend[0].path.endswith("cpusim.so")
    rng = np.random.default_rng(12)
    buf = bytearray(rng.integers(0, 256, 400_000 if small else 900_000, dtype=np.uint8).tobytes())
    pos = 3000
    while pos < len(buf) - 400:  # copies of what lies a little further back, at random distances: a few hundred sequences per block
        n = int(rng.integers(8, 200))
        back = int(rng.integers(n, 2500))
        buf[pos:pos + n] = buf[pos - back:pos - back + n]
        pos += n + int(rng.integers(20, 600))
    data = bytes(buf)
    sig = ol.stage1_encode(data_gen.nanopore_like(150_000 if small else 600_000, rng), 2, True)
    frames, sizes, want = [], [], []
    for level in (1, 7, 19):
        for d in (data, sig, data[:150_000]):
            frames.append(ol.zstd_compress_lib(d, level))
            sizes.append(len(d))
            want.append(d)
    modes = set()
    for f in frames:  # (the scenario is there: some block repeats one table, some block all three, some none)
        p = 4 + 1 + (0, 1, 2, 4)[f[4] & 3] + ((1 if (f[4] >> 5) & 1 else 0, 2, 4, 8)[f[4] >> 6]) + (0 if (f[4] >> 5) & 1 else 1)
        while True:
            bh = int.from_bytes(f[p:p + 3], "little")
            p += 3
            if (bh >> 1) & 3 == 2:
                b0 = f[p]
                lt, sf = b0 & 3, (b0 >> 2) & 3
                if lt < 2:
                    hdr, regen = (1, b0 >> 3) if sf in (0, 2) else ((2, int.from_bytes(f[p:p + 2], "little") >> 4) if sf == 1 else (3, int.from_bytes(f[p:p + 3], "little") >> 4))
                    sec = hdr + (regen if lt == 0 else 1)
                else:
                    v = int.from_bytes(f[p:p + 5], "little")
                    hdr, comp = ((3, (v >> 14) & 0x3FF) if sf in (0, 1) else ((4, (v >> 18) & 0x3FFF) if sf == 2 else (5, (v >> 22) & 0x3FFFF)))
                    sec = hdr + comp
                n0 = f[p + sec]
                if n0:
                    modes.add(f[p + sec + (1 if n0 < 128 else 2 if n0 < 255 else 3)])
            p += 1 if (bh >> 1) & 3 == 1 else bh >> 3
            if bh & 1:
                break
    rep = lambda m: [(m >> s) & 3 == 3 for s in (6, 4, 2)]
    assert any(all(rep(m)) for m in modes) and any(any(rep(m)) and not all(rep(m)) for m in modes) and any(not any(rep(m)) for m in modes), modes
    res = batch.decompress(frames, sizes, Options(False, 0, 1, 0), sized=False)
    for i, (d, (code, blob)) in enumerate(zip(want, res)):
        assert code == len(d) and blob == d, (i, code, len(d))
    st = _decode_stats(backend)
    assert st["general"] == 0 and st["seq_frames"] == len(frames), st


def test_frames_that_run_the_pool_dry_leave_nothing_behind_for_their_neighbours(backend):
    """Unit slots are allotted before the walker knows whether a frame's records fit the pool; a frame that does not fit leaves its slots
    unwritten.  They must not keep what an earlier batch of the same context left there: the executor follows a frame's Seq units by
    their read index, and a valid frame in front of such slots was decoded into the error code (found by a fuzz run over batches)."""
    _, batch, _ = backend
    rng = np.random.default_rng(31)
    words = [bytes(rng.integers(97, 123, int(rng.integers(2, 9)), dtype=np.uint8)) for _ in range(200)]

    def text(n):  # a sequence every few bytes: more records than the pool holds for a frame (one per 32 bytes)
        out, tot = [], 0
        while tot < n:
            w = words[int(rng.integers(0, 200))]
            out.append(w + b" ")
            tot += len(w) + 1
        return b"".join(out)[:n]

    def copies(n):  # noise with a copy of 100 - 250 bytes every few hundred: a sequence per ~400 bytes
        buf = bytearray(rng.integers(0, 256, n, dtype=np.uint8).tobytes())
        pos = 3000
        while pos < n - 600:
            m = int(rng.integers(100, 250))
            back = int(rng.integers(m, 2500))
            buf[pos:pos + m] = buf[pos - back:pos - back + m]
            pos += m + int(rng.integers(100, 400))
        return bytes(buf)

    o = Options(False, 0, 1, 0)
    # batch A: six frames of five blocks with sequences each - the unit slots now hold Seq units of reads 0 .. 5
    a = [copies(600_000) for _ in range(6)]
    res = batch.decompress([ol.zstd_compress_lib(d, 3) for d in a], [len(d) for d in a], o, sized=False)
    assert all(code == len(d) and blob == d for d, (code, blob) in zip(a, res))
    st = _decode_stats(backend)
    assert st["general"] == 0 and st["seq_frames"] == 6, st
    # batch B: a short frame that fits, in front of one that runs the pool dry (45 k sequences, 14 k records for the two)
    b = [copies(100_000), text(350_000)]
    for _ in range(2):
        res = batch.decompress([ol.zstd_compress_lib(d, 1) for d in b], [len(d) for d in b], o, sized=False)
        for i, (d, (code, blob)) in enumerate(zip(b, res)):
            assert code == len(d) and blob == d, (i, hex(code))
        st = _decode_stats(backend)
        assert st["general"] >= 1 and st["seq_frames"] >= 1, st  # (the scenario: some frames ran the pool dry, the first did not)
