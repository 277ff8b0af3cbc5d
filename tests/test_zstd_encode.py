"""Stage 2, encode direction: every frame the product writes (csrc/zstd_enc_kernels.cuh) must decode with stock
libzstd and with the reference / oracle to the original bytes, must carry its content size, and should be about as
small as what libzstd level 1 writes for nanopore-like signal (the ratio is reported by bench.py; here only a loose
bound is asserted).  Runs on the CUDA emulator build (CPU suite) and on the real library (-m gpu)."""
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


def _check_frame(frame, content):
    assert ol.frame_content_size(frame) == len(content)          # vbz.cpp:236-240 needs it
    assert ol.zstd_decompress_lib(frame, len(content)) == content  # stock libzstd
    assert ol.zstd_decompress_port(frame, len(content)) == content  # the restated decoder


def _payloads(rng, small):
    yield b""
    yield b"\\x07"
    yield bytes([5, 4, 3, 2, 1, 9, 8])
    yield bytes(100)
    yield bytes(rng.integers(0, 256, 1000, dtype=np.uint8))
    yield bytes(rng.integers(0, 3, 5000, dtype=np.uint8))
    yield bytes(rng.integers(0, 200, 40000, dtype=np.uint8) // 3)
    yield ol.stage1_encode(data_gen.nanopore_like(20000, rng), 2, True)
    yield bytes((rng.normal(128, 4, 70000)).astype(np.uint8))              # a few very frequent and many very rare symbols: long codes
    if not small:
        yield ol.stage1_encode(data_gen.nanopore_like(300000, rng), 2, True)
        yield bytes(rng.integers(0, 256, 100000, dtype=np.uint8))


def test_zstd_only_frames_are_valid(backend):
    capi = backend[0]
    rng = np.random.default_rng(2)
    small = capi.path.endswith("cpusim.so")
    for data in _payloads(rng, small):
        for level in (1, 3, 100):
            o = Options(False, 0, level, 0)
            r, frame = capi.compress(np.frombuffer(data, np.uint8), o, sized=False)
            assert not ol.is_error(r), (len(data), level, r)
            _check_frame(frame, data)
            r2, back = capi.decompress(frame, o, len(data), sized=False)   # and our own decoder
            assert r2 == len(data) and back == data
        # a destination that is too small is a zstd error (vbz.cpp:201-204)
        if len(data) > 2000:
            assert capi.compress(np.frombuffer(data, np.uint8), Options(False, 0, 1, 0), sized=False, cap=64)[0] == ol.ERR_ZSTD


def test_kat_tiny_input_is_a_raw_block(backend):
    """vbz_test.cpp:238: {5,4,3,2,1} int32 zigzag level 100 -> a frame holding one raw block."""
    capi = backend[0]
    k = [k for k in golden_data.kats() if k["lvl"] != 0 and k["isz"] == 4][0]
    arr = np.array(k["values"], dtype=k["dtype"])
    r, out = capi.compress(arr, Options(k["zz"], k["isz"], k["lvl"], k["ver"]), sized=k["sized"])
    assert list(out) == k["bytes"], k["source"]


@pytest.mark.parametrize("dtype,o", [(np.int16, Options(True, 2, 1, 0)), (np.int16, Options(True, 2, 1, 1)), (np.int16, Options(False, 2, 3, 0)),
                                     (np.int32, Options(True, 4, 1, 0)), (np.uint32, Options(False, 4, 3, 0)), (np.int8, Options(True, 1, 1, 0)), (np.int8, Options(True, 1, 1, 1)), (np.uint8, Options(False, 1, 2, 1))])
def test_full_vbz_roundtrip_and_reference_decodes(backend, dtype, o):
    capi = backend[0]
    ref = ol.ref()
    rng = np.random.default_rng(7)
    small = capi.path.endswith("cpusim.so")
    for n in (0, 1, 7, 100, 4096, 20000, 150000) + ((400001,) if o.integer_size == 2 else ()):  # up to several 128 KiB table groups
        if dtype == np.uint32:
            arr = rng.integers(0, 1 << 20, n).astype(np.uint32)
        else:
            arr = data_gen.nanopore_like(n, rng).astype(dtype)
        for sized in (True, False):
            r, blob = capi.compress(arr, o, sized=sized)
            assert not ol.is_error(r), (n, sized, r)
            # the oracle (and the reference itself, where it was built) decode it to the original samples
            rp, back = port.decompress(blob, o, arr.nbytes, sized=sized)
            assert rp == arr.nbytes and back == arr.tobytes(), (n, sized, rp)
            if ref is not None:
                rr, back2 = ref.decompress(blob, o, arr.nbytes, sized=sized)
                assert rr == arr.nbytes and back2 == arr.tobytes()
            # stage-1 bytes inside the frame are the oracle's stage-1 bytes
            frame = blob[4:] if sized else blob
            s1 = ol.stage1_encode(arr, o.integer_size, o.perform_delta_zig_zag, o.vbz_version)
            assert ol.zstd_decompress_lib(frame, len(s1)) == s1
            # our decoder
            r3, back3 = capi.decompress(blob, o, arr.nbytes, sized=sized)
            assert r3 == arr.nbytes and back3 == arr.tobytes()


def test_groups_raw_rle_and_mixed_content(backend):
    """Table groups of every kind: incompressible (one Raw block per 128 KiB), constant (one RLE block), and content whose
    statistics change inside a group (sub-blocks fall back to Raw one by one; the tree travels with the first Huffman one)."""
    capi = backend[0]
    rng = np.random.default_rng(17)
    o = Options(False, 0, 1, 0)
    cases = [rng.integers(0, 256, 300000).astype(np.uint8), np.zeros(400000, np.uint8), np.full(131072, 7, np.uint8),
             np.concatenate([np.zeros(200000, np.uint8), rng.integers(0, 256, 100000).astype(np.uint8), rng.integers(0, 4, 200000).astype(np.uint8)]),
             np.concatenate([rng.integers(0, 256, 9000).astype(np.uint8), rng.integers(0, 3, 120000).astype(np.uint8)]),
             rng.integers(0, 5, 8192 * 3 + 1023).astype(np.uint8), rng.integers(0, 5, 8192 + 31).astype(np.uint8), rng.integers(0, 5, 1024).astype(np.uint8)]
    for arr in cases:
        for sized in (True, False):
            r, blob = capi.compress(arr, o, sized=sized)
            assert not ol.is_error(r)
            frame = blob[4:] if sized else blob
            assert ol.zstd_decompress_lib(frame, arr.nbytes) == arr.tobytes()
            r3, back3 = capi.decompress(blob, o, arr.nbytes, sized=sized)
            assert r3 == arr.nbytes and back3 == arr.tobytes()
    # a constant group is one RLE block; random bytes cost 3 bytes per 128 KiB on top of the frame header
    assert capi.compress(cases[1], o, sized=False)[0] < 64
    assert capi.compress(cases[0], o, sized=False)[0] <= 300000 + 9 + 3 * 3


def test_ratio_close_to_libzstd_on_signal(backend):
    capi = backend[0]
    rng = np.random.default_rng(12)
    arr = data_gen.nanopore_like(120000, rng)
    o = Options(True, 2, 1, 1)
    r, blob = capi.compress(arr, o, sized=True)
    rp, blob_ref = port.compress(arr, o, sized=True)
    assert not ol.is_error(r)
    assert r <= 1.01 * rp, (r, rp)   # zero runs of the control bytes as sequences, Huffman literals: what libzstd level 1 reaches on signal


def test_fast5_fixture_reencode(backend):
    """BASELINE config 1 on the product: re-encode the fixture reads, decode with the oracle, compare sizes."""
    capi = backend[0]
    o = Options(True, 2, 1, 1)
    tot, tot_ref = 0, 0
    reads = golden_data.fast5_reads()
    for rd in (reads[:3] if capi.path.endswith("cpusim.so") else reads):
        want = rd["samples"]
        r, blob = capi.compress(want, o, sized=True)
        assert not ol.is_error(r)
        rp, back = port.decompress(blob, o, want.nbytes, sized=True)
        assert rp == want.nbytes and back == want.tobytes()
        tot += r
        tot_ref += len(rd["vbz_chunk"])
    assert tot <= 1.006 * tot_ref, (tot, tot_ref)  # the reference's own files: 0.4222 of raw; this writer 0.4231


def test_repetitive_inputs_compress_like_libzstd(backend):
    """Where matches are everything (a Huffman-only writer was 7-90x worse here): the pyvbz README's arange (27 bytes with the
    reference), the reference benchmark's tiled trace (vbz/perf/test_data_generator.h:29-74), constant signal - sizes next to what the
    reference writes at the same level, and every frame decodes with the reference."""
    capi = backend[0]
    ref = ol.ref() or port
    rng = np.random.default_rng(3)
    small = capi.path.endswith("cpusim.so")
    trace = data_gen.nanopore_like(2000, rng)
    fixture = golden_data.fast5_reads()[0]["samples"]
    cases = [("arange", np.arange(1000, dtype=np.int16), 40, None), ("tiled synthetic trace", np.tile(trace, 30 if small else 100), None, 1.5),
             ("tiled fixture trace", np.tile(fixture[:5000], 10 if small else 30), None, 1.5), ("constant", np.full(60000, 7, np.int16), 64, None),
             # a period as long as the reference benchmark's trace (15,643 samples = 16 KB of data bytes: the look-back into the group before)
             ("long period", np.resize(data_gen.nanopore_like(15643, rng), 90000 if small else 190000), None, 1.15),
             ("zeros", np.zeros(100000, np.int16), 64, None)]
    for level in (1, 3):
        o = Options(True, 2, level, 1)
        for name, arr, max_bytes, max_factor in cases:
            r, blob = capi.compress(arr, o, sized=True)
            rr, theirs = ref.compress(arr, o, sized=True)
            assert not ol.is_error(r) and not ol.is_error(rr)
            if max_bytes is not None:
                assert r <= max_bytes, (name, level, r, rr)
            if max_factor is not None:
                assert r <= max_factor * rr, (name, level, r, rr)
            rd, back = ref.decompress(blob, o, arr.nbytes, sized=True)
            assert rd == arr.nbytes and back == arr.tobytes(), (name, level)
            r3, back3 = capi.decompress(blob, o, arr.nbytes, sized=True)
            assert r3 == arr.nbytes and back3 == arr.tobytes(), (name, level)


def test_runs_of_every_length_among_control_bytes(backend):
    """The run scan of the control-byte groups (csrc/zstd_match_kernels.cuh: zm_runs_line / zm_runs_bytes): flat stretches of every length and
    alignment - shorter than a lane, whole 128-byte lines of zero control bytes, runs across lines and spans - between bursts of large
    deltas.  A run that is lost or cut short still decodes, so the size is the check: next to what the reference writes."""
    capi = backend[0]
    ref = ol.ref() or port
    rng = np.random.default_rng(5)
    small = capi.path.endswith("cpusim.so")
    for trial in range(2 if small else 6):
        pieces, level = [], 100
        for L in rng.integers(8, 9000, 40 if small else 120):
            pieces.append(np.full(int(L), level, np.int16))
            burst = rng.integers(-3000, 3000, int(rng.integers(1, 40))).astype(np.int16)
            pieces.append(burst)
            level = int(rng.integers(-500, 500))
        arr = np.concatenate(pieces)
        o = Options(True, 2, 1, 1)
        r, blob = capi.compress(arr, o, sized=True)
        rr, theirs = ref.compress(arr, o, sized=True)
        assert not ol.is_error(r) and not ol.is_error(rr)
        assert r <= 1.10 * rr + 16, (trial, r, rr)
        rd, back = ref.decompress(blob, o, arr.nbytes, sized=True)
        assert rd == arr.nbytes and back == arr.tobytes()
        r3, back3 = capi.decompress(blob, o, arr.nbytes, sized=True)
        assert r3 == arr.nbytes and back3 == arr.tobytes()


def test_matches_of_every_kind_round_trip(backend):
    """The match finder on text-like and structured bytes (zstd only, every level the reference maps to a strategy of its own): long and
    short matches, overlapping ones, matches across the 128 KiB groups and the 8 KiB literal blocks inside them; stock libzstd decodes all."""
    capi = backend[0]
    rng = np.random.default_rng(23)
    small = capi.path.endswith("cpusim.so")
    words = [bytes(rng.integers(97, 123, int(k), dtype=np.uint8)) for k in rng.integers(2, 12, 300)]
    text = b" ".join(words[int(i)] for i in rng.integers(0, 300, 20000 if small else 80000))
    cases = [text, b"ab" * 70000, bytes(rng.integers(0, 256, 5000, dtype=np.uint8)) * (20 if small else 60),
             bytes(np.arange(50000 if small else 150000, dtype=np.uint32).view(np.uint8)), bytes(200000) + b"x" + bytes(70000),
             bytes(np.repeat(rng.integers(0, 256, 3000, dtype=np.uint8), rng.integers(1, 40, 3000)))]
    for level in (1, 2, 3, 19):
        o = Options(False, 0, level, 0)
        for data in cases:
            arr = np.frombuffer(data, np.uint8)
            r, frame = capi.compress(arr, o, sized=False)
            assert not ol.is_error(r), (level, len(data), r)
            _check_frame(frame, data)
            r2, back = capi.decompress(frame, o, len(data), sized=False)
            assert r2 == len(data) and back == data
            assert r <= len(ol.zstd_compress_lib(data, 1)) * 2 + 64, (level, len(data), r)  # within 2x of libzstd level 1 even on text
