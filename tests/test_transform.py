"""Stage 1 (delta + zigzag + StreamVByte, zstd level 0) through the C-ABI of the product library,
compared byte for byte with the oracle.  Runs twice: on the CUDA emulator build (CPU suite) and on
the real library on a B200 (-m gpu)."""
import ctypes

import numpy as np
import pytest

import data_gen
import golden_data
import oracle_lib as ol
import product
from oracle_lib import Options
from vbz_compression_b200 import CompressionOptions

port = ol.port()


@pytest.fixture(params=product.BACKENDS)
def backend(request):
    return product.get(request.param)


LEVEL0_OPTS = [Options(zz, isz, 0, ver) for zz in (True, False) for isz in (1, 2, 4) for ver in (0, 1)]  # incl. the v1 "half" packing of int8


@pytest.mark.parametrize("k", [k for k in golden_data.kats() if k["lvl"] == 0], ids=lambda k: k["name"])
def test_kat_bytes(backend, k):
    capi = backend[0]
    arr = np.array(k["values"], dtype=k["dtype"])
    o = Options(k["zz"], k["isz"], k["lvl"], k["ver"])
    r, out = capi.compress(arr, o, sized=k["sized"])
    assert not ol.is_error(r) and list(out) == k["bytes"], k["source"]
    r2, back = capi.decompress(bytes(k["bytes"]), o, arr.nbytes, sized=k["sized"])
    assert r2 == arr.nbytes and back == arr.tobytes()


def _cases(rng):
    yield np.zeros(0, np.int16)
    for n in (1, 3, 4, 5, 8, 9, 31, 33, 100, 255, 256, 257, 511, 512, 513, 1024, 4095, 4096, 4097, 8192, 9000):
        yield data_gen.nanopore_like(n, rng)
        yield data_gen.uniform_full_range(n, np.int16, rng)
    for n in (1, 7, 2047, 2048, 2049, 5000):
        yield data_gen.uniform_full_range(n, np.int32, rng)
        yield (rng.integers(0, 1 << 20, n)).astype(np.uint32)
        yield (rng.integers(-300, 300, n)).astype(np.int32)
    for n in (1, 15, 16, 17, 511, 512, 513, 1025, 8191, 8192, 8193, 20000):
        yield data_gen.uniform_full_range(n, np.int8, rng)
        yield (rng.integers(-3, 4, n)).astype(np.int8)
        yield (rng.integers(-1, 2, n) * (rng.random(n) < 0.3)).astype(np.int8)  # many zero values: code 0, no data nibbles
    yield data_gen.nanopore_like(70_000, rng)


def test_single_calls_match_oracle(backend):
    capi = backend[0]
    rng = np.random.default_rng(3)
    for arr in _cases(rng):
        raw = arr.view(np.uint8)
        for o in LEVEL0_OPTS:
            for sized in (True, False):
                rp, bp = port.compress(raw, o, sized=sized)
                rg, bg = capi.compress(raw, o, sized=sized)
                assert rg == rp, (arr.dtype, arr.size, o.integer_size, o.perform_delta_zig_zag, sized, rg, rp)
                if ol.is_error(rp) or (sized and rp < 4):
                    continue
                assert bg == bp, (arr.dtype, arr.size, o.integer_size, o.perform_delta_zig_zag, sized)
                r2, d2 = capi.decompress(bp, o, raw.size, sized=sized)
                assert r2 == raw.size and d2 == raw.tobytes(), (arr.dtype, arr.size, o.integer_size, o.perform_delta_zig_zag, sized, r2)


def test_fast5_fixture_stage1(backend):
    """The stage-1 bytes inside the reference's stored frames must equal our encoder's output for the
    decoded samples, and decode back."""
    capi = backend[0]
    o = Options(True, 2, 0, 0)
    for rd in golden_data.fast5_reads()[:6]:
        want = rd["samples"]
        s1 = ol.zstd_decompress_lib(rd["vbz_chunk"][4:], ol.frame_content_size(rd["vbz_chunk"][4:]))
        r, got = capi.compress(want, o, sized=False)
        assert r == len(s1) and got == s1
        r2, back = capi.decompress(s1, o, want.nbytes, sized=False)
        assert r2 == want.nbytes and back == want.tobytes()


def test_batch_matches_single_and_isolates_bad_reads(backend):
    _, batch, _ = backend
    rng = np.random.default_rng(5)
    o = Options(True, 2, 0, 0)
    reads = [data_gen.nanopore_like(int(n), rng) for n in (10, 0, 5000, 4096, 12345, 1, 8191)]
    res = batch.compress(reads, o, sized=True)
    for arr, (code, blob) in zip(reads, res):
        rp, bp = port.compress(arr, o, sized=True)
        assert code == rp and blob == bp
    # decode the batch with one corrupted and one truncated member
    chunks = [b for _, b in res]
    chunks[2] = chunks[2][:-3]
    bad = bytearray(chunks[4]); bad[4] ^= 0xFF; chunks[4] = bytes(bad)
    sizes = [a.nbytes for a in reads]
    out = batch.decompress(chunks, sizes, o, sized=True)
    for i, (arr, (code, blob)) in enumerate(zip(reads, out)):
        rp, bp = port.decompress(chunks[i], o, arr.nbytes, sized=True)
        assert code == rp, (i, code, rp)
        if not ol.is_error(rp):
            assert blob == bp
    assert out[0][1] == reads[0].tobytes() and out[6][1] == reads[6].tobytes()
    assert ol.is_error(out[2][0])


def test_int16_decoder_on_foreign_and_damaged_streams(backend):
    """The int16 path's scalar branch: streams with 3-byte codes (what a non-x86 build of the reference emits for
    |delta| > 32767), streams cut short and streams with trailing bytes, over several
    tiles so that offsets and the running sample value cross tile borders.  Checked against the oracle in x86 mode."""
    capi = backend[0]
    rng = np.random.default_rng(11)
    o = Options(True, 2, 0, 0)
    for n in (37, 9000, 20011):
        arr = data_gen.uniform_full_range(n, np.int16, rng)
        arr[n // 2:] = data_gen.nanopore_like(n - n // 2, rng)  # wide codes in the first half only
        port.lib.vbzo_set_int16_x86_semantics(0)
        try:
            r, wide = port.compress(arr, o, sized=False)  # generic template: 32-bit deltas
        finally:
            port.lib.vbzo_set_int16_x86_semantics(1)
        assert not ol.is_error(r)
        _, narrow = port.compress(arr, o, sized=False)
        assert len(wide) > len(narrow)
        for blob in (wide, wide[:-1], wide + b"\x01", narrow[:-5], narrow[:len(narrow) // 2], narrow + b"\0\0"):
            a = capi.decompress(blob, o, arr.nbytes, sized=False)
            b = port.decompress(blob, o, arr.nbytes, sized=False)
            assert a[0] == b[0], (n, len(blob), a[0], b[0])
            if not ol.is_error(b[0]):
                assert a[1] == b[1], (n, len(blob))


def test_error_paths_match_oracle(backend):
    capi = backend[0]
    arr = np.arange(100, dtype=np.int16)
    cases = [
        (Options(True, 3, 0, 0), True), (Options(True, 2, 0, 2), True), (Options(True, 2, 0, 2), False),
        (Options(True, 4, 0, 0), True),   # 200 % 4 == 0 fine
        (Options(True, 0, 0, 0), True), (Options(False, 0, 0, 1), False),
    ]
    for o, sized in cases:
        assert capi.compress(arr, o, sized=sized) == port.compress(arr, o, sized=sized)
    odd = np.zeros(7, np.uint8)
    for isz in (2, 4):
        for sized in (True, False):
            o = Options(True, isz, 0, 0)
            assert capi.compress(odd, o, sized=sized, cap=128)[0] == port.compress(odd, o, sized=sized, cap=128)[0]
    # capacity below the reference's requirement at level 0 (vbz.cpp:171-174)
    o = Options(True, 2, 0, 0)
    assert capi.compress(arr, o, sized=False, cap=100)[0] == port.compress(arr, o, sized=False, cap=100)[0] == ol.ERR_DEST_SIZE
    # decode-side
    _, good = port.compress(arr, o, sized=True)
    for blob, ds in ((good, 200), (good, 199), (good, 202), (good[:3], 200), (good[:-1], 200), (good + b"\0", 200), (b"", 0), (b"\0\0\0\0", 0)):
        for sized in (True, False):
            for oo in (o, Options(True, 4, 0, 0), Options(False, 2, 0, 0), Options(True, 1, 0, 0), Options(True, 1, 0, 1), Options(False, 1, 0, 1), Options(True, 2, 0, 3), Options(True, 0, 0, 0)):
                a = capi.decompress(blob, oo, ds, sized=sized)
                b = port.decompress(blob, oo, ds, sized=sized)
                assert a == b, (len(blob), ds, sized, oo.integer_size, oo.vbz_version, a[0], b[0])


def test_fuzz_corpus_level0(backend):
    """Restates vbz/fuzzing/vbz_fuzz.cpp for the option combos served so far (level 0, v0 packing)."""
    capi = backend[0]
    corpus = golden_data.fuzz_corpus()
    step = 1 if backend[2] is not None and not backend[0].path.endswith("cpusim.so") else 6
    for blob in corpus[::step]:
        n = len(blob)
        for o in LEVEL0_OPTS[:: (1 if step == 1 else 2)]:
            for sized in (True, False):
                assert capi.compress(blob, o, sized=sized) == port.compress(blob, o, sized=sized)
                for ds in sorted(set([0, 2, 4, 8, 16, 32, 64, n, 2 * n, max(4 * (n - 4), 0)])):
                    a = capi.decompress(blob, o, ds, sized=sized)
                    b = port.decompress(blob, o, ds, sized=sized)
                    assert a[0] == b[0], (n, ds, sized, o.integer_size, o.perform_delta_zig_zag, a[0], b[0])
                    if not ol.is_error(b[0]):
                        assert a[1] == b[1]


def test_library_exports_every_declared_symbol():
    from vbz_compression_b200 import _lib
    import os
    path = product.vbuild.build_cpusim()
    lib = ctypes.CDLL(path)
    for s in _lib.EXPORTED_SYMBOLS:
        assert hasattr(lib, s), s
    native = _lib.LIB_PATH
    if os.path.exists(native):
        import subprocess
        syms = subprocess.run(["nm", "-D", "--defined-only", native], capture_output=True, text=True, check=True).stdout
        for s in _lib.EXPORTED_SYMBOLS:
            assert (" T " + s) in syms, s


@pytest.mark.parametrize("kind", product.BACKENDS)
def test_reads_longer_than_the_prefetched_lookback(kind):
    """A tile's look-back walks back to the first tile of its read; four windows of 32 tiles are prefetched into shared
    memory (csrc/svb_kernels.cuh: lookback_exclusive_win), longer walks continue with direct loads.  A read of more than
    128 tiles (> 1,048,576 samples) next to short ones must still come out byte-exact, both ways."""
    capi, batch, _ = product.get(kind)
    rng = np.random.default_rng(314)
    long_n = 1_250_000 if kind == "cpusim" else 5_000_000
    reads = [data_gen.nanopore_like(n, rng) for n in (30_000, long_n, 9, 70_000)]
    o = Options(True, 2, 0, 0)
    opts = CompressionOptions(True, 2, 0, 0)
    res = batch.compress(reads, opts, sized=True)
    for arr, (code, blob) in zip(reads, res):
        rp, bp = port.compress(arr, o, sized=True)
        assert code == rp and blob == bp
    back = batch.decompress([b for _, b in res], [a.nbytes for a in reads], opts, sized=True)
    for arr, (code, blob) in zip(reads, back):
        assert code == arr.nbytes and blob == arr.tobytes()
