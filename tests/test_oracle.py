"""CPU tests (-m "not gpu"): pin the oracle port (oracle/vbz_oracle.c, zstd_oracle.c) against
  1. the reference's byte-exact known-answer vectors            (tests/golden/kat.json)
  2. the reference's stored fast5 fixtures                      (tests/golden/fast5_chunks.npz)
  3. the compiled reference itself, oracle/_ref                 (random inputs, every option combo)
  4. the reference's fuzz corpus                                (tests/golden/fuzz_corpus.npz)
  5. libzstd 1.5.5 frames at many levels for the restated zstd decoder."""
import itertools
import struct

import numpy as np
import pytest

import data_gen
import golden_data
import oracle_lib as ol
from oracle_lib import Options

port = ol.port()
ref = ol.ref()
needs_ref = pytest.mark.skipif(ref is None, reason="oracle/_ref not built (no /root/reference and no prebuilt copy)")

ALL_OPTS = [Options(zz, isz, lvl, ver) for zz in (True, False) for isz in (0, 1, 2, 4) for lvl in (0, 1) for ver in (0, 1)]


def _o(k):
    return Options(k["zz"], k["isz"], k["lvl"], k["ver"])


@pytest.mark.parametrize("k", golden_data.kats(), ids=lambda k: k["name"])
def test_kat_bytes(k):
    arr = np.array(k["values"], dtype=k["dtype"])
    r, out = port.compress(arr, _o(k), sized=k["sized"])
    assert not ol.is_error(r)
    if k["lvl"] == 0 or len(k["values"]) <= 8:  # zstd bytes are only pinned for raw-block frames
        assert list(out) == k["bytes"], k["source"]
    r2, back = port.decompress(bytes(k["bytes"]), _o(k), arr.nbytes, sized=k["sized"])
    assert r2 == arr.nbytes and back == arr.tobytes()


def test_fast5_fixtures_decode_and_stage1_bytes():
    total_raw = total_vbz = 0
    for rd in golden_data.fast5_reads():
        chunk, want = rd["vbz_chunk"], rd["samples"]
        assert rd["v1_identical"]
        for ver in (0, 1):
            o = Options(True, 2, 1, ver)
            assert port.decompressed_size(chunk, o) == want.nbytes
            r, got = port.decompress(chunk, o, want.nbytes, sized=True)
            assert r == want.nbytes and got == want.tobytes()
        # restated zstd decoder == libzstd on the stored frame, and stage-1 bytes == port's encode
        fcs = ol.frame_content_size(chunk[4:])
        s1 = ol.zstd_decompress_port(chunk[4:], fcs)
        assert s1 is not None and s1 == ol.zstd_decompress_lib(chunk[4:], fcs)
        assert s1 == ol.stage1_encode(want, 2, True)
        total_raw += want.nbytes
        total_vbz += len(chunk)
    assert (total_raw, total_vbz) == (3097862, 1308058)


def test_error_strings_and_predicates():
    names = {ol.ERR_ZSTD: "VBZ_ZSTD_ERROR", ol.ERR_INPUT_SIZE: "VBZ_INPUT_SIZE_ERROR", ol.ERR_INTEGER_SIZE: "VBZ_INTEGER_SIZE_ERROR",
             ol.ERR_DEST_SIZE: "VBZ_DESTINATION_SIZE_ERROR", ol.ERR_STREAM: "VBZ_STREAMVBYTE_STREAM_ERROR",
             ol.ERR_VERSION: "VBZ_VERSION_ERROR", ol.ERR_OOM: "VBZ_OUT_OF_MEMORY_ERROR", 5: "VBZ_UNKNOWN_ERROR", 0xFFFFFFF8: "VBZ_UNKNOWN_ERROR"}
    for v, s in names.items():
        assert port.error_string(v) == s
        assert port.is_error(v) == (v >= 0xFFFFFFF9)
        if ref:
            assert ref.error_string(v) == s and ref.is_error(v) == port.is_error(v)


@needs_ref
def test_max_compressed_size_matches_reference():
    for o in ALL_OPTS + [Options(True, 3, 1, 0), Options(True, 2, 1, 2), Options(False, 4, 3, 7)]:
        for n in (0, 1, 2, 3, 4, 7, 8, 100, 1000, 131071, 131072, 2_000_000, 0x3FFFFFFC, 0xFFFFFFF0):
            assert port.max_compressed_size(n, o) == ref.max_compressed_size(n, o), (n, o.integer_size, o.zstd_compression_level)
    assert port.max_compressed_size(2_000_000, Options(True, 2, 1, 1)) == 4266621
    assert port.max_compressed_size(2_000_000, Options(True, 2, 0, 0)) == 4250020


def _inputs(rng):
    yield np.zeros(0, np.int16)
    for n in (1, 2, 3, 4, 5, 7, 8, 9, 15, 16, 17, 31, 32, 33, 63, 64, 65, 100, 1000, 4099):
        yield data_gen.nanopore_like(n, rng)
        yield data_gen.uniform_full_range(n, np.int16, rng)
        yield data_gen.uniform_full_range(n, np.int32, rng)
        yield data_gen.uniform_full_range(n, np.int8, rng)
        yield (rng.integers(0, 1 << 20, n)).astype(np.uint32)
        yield np.arange(n, dtype=np.int16)
    yield data_gen.nanopore_like(50_000, rng)
    yield data_gen.uniform_full_range(30_000, np.int16, rng)


@needs_ref
@pytest.mark.parametrize("generic", [False, True], ids=["x86", "generic"])
def test_port_equals_reference_on_random_inputs(generic):
    """Level-0 bytes must be identical; level-1 frames must cross-decode; every decode must agree."""
    rlib = ol.ref(generic=generic)
    port.lib.vbzo_set_int16_x86_semantics(0 if generic else 1)
    try:
        rng = np.random.default_rng(7)
        for arr in _inputs(rng):
            raw = arr.view(np.uint8)
            for o in ALL_OPTS:
                for sized in (True, False):
                    rp, bp = port.compress(raw, o, sized=sized)
                    rr, br = rlib.compress(raw, o, sized=sized)
                    assert rp == rr, (arr.dtype, arr.size, o.integer_size, o.zstd_compression_level, rp, rr)
                    if ol.is_error(rr) or (sized and rr < 4):
                        continue
                    assert bp == br
                    r1, d1 = port.decompress(br, o, raw.size, sized=sized)
                    r2, d2 = rlib.decompress(bp, o, raw.size, sized=sized)
                    assert r1 == r2 == raw.size and d1 == d2 == raw.tobytes()
    finally:
        port.lib.vbzo_set_int16_x86_semantics(1)


def _legacy_zstd(payload):
    """zstd v0.2-v0.7 frame magics (0xFD2FB522..27).  The image's libzstd is built with
    ZSTD_LEGACY_SUPPORT and decodes them; no VBZ writer can produce them (VBZ needs zstd >= 1.3.1,
    CMakeLists.txt:156), so the port and the product reject them with VBZ_ZSTD_ERROR - the one
    documented divergence on garbage input ('out of scope')."""
    return len(payload) >= 4 and payload[1:4] == b"\xb5\x2f\xfd" and 0x22 <= payload[0] <= 0x27


@needs_ref
def test_port_equals_reference_on_fuzz_corpus():
    """Restates vbz/fuzzing/vbz_fuzz.cpp:165-197: every corpus file x every option combo, compress
    round trip and decode-of-garbage at candidate sizes; port and reference must return the same
    code and, on success, the same bytes."""
    corpus = golden_data.fuzz_corpus()
    assert len(corpus) == 238
    mism = []
    for ci, blob in enumerate(corpus):
        for o in ALL_OPTS:
            for sized in (True, False):
                rp, bp = port.compress(blob, o, sized=sized)
                rr, br = ref.compress(blob, o, sized=sized)
                if rp != rr or bp != br:
                    mism.append(("c", ci, o.perform_delta_zig_zag, o.integer_size, o.zstd_compression_level, o.vbz_version, sized, rp, rr))
            # decode garbage: a spread of candidate destination sizes (the fuzzer sweeps them all)
            n = len(blob)
            cands = sorted(set([0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 12, 15, 16, 17, 24, 31, 32, 33, 40, 48, 64, 100, 128, 255, 256, n, 2 * n, 4 * n, max(n - 1, 0),
                                max(n - 4, 0), max(4 * (n - 4), 0), 1000, 4096]))
            for ds in cands:
                for sized in (True, False):
                    r1, d1 = port.decompress(blob, o, ds, sized=sized)
                    r2, d2 = ref.decompress(blob, o, ds, sized=sized)
                    if o.zstd_compression_level and _legacy_zstd(blob[4:] if sized else blob):
                        assert r1 == ol.ERR_ZSTD or (r1, d1) == (r2, d2)
                        continue
                    if r1 != r2 or d1 != d2:
                        mism.append(("d", ci, o.perform_delta_zig_zag, o.integer_size, o.zstd_compression_level, o.vbz_version, sized, ds, r1, r2))
    assert not mism, (len(mism), mism[:20])


def _zstd_inputs(rng):
    yield b""
    yield b"a"
    yield b"abc" * 5
    yield bytes(1000)
    yield bytes(rng.integers(0, 256, 100, dtype=np.uint8))
    yield bytes(rng.integers(0, 256, 200_000, dtype=np.uint8))
    yield bytes(rng.integers(0, 4, 50_000, dtype=np.uint8))
    yield bytes(rng.integers(0, 16, 300_000, dtype=np.uint8))
    yield (b"the quick brown fox jumps over the lazy dog " * 5000)[:150_001]
    yield np.arange(1000, dtype=np.int16).tobytes()
    yield ol.stage1_encode(data_gen.nanopore_like(400_000, rng), 2, True)
    yield ol.stage1_encode(data_gen.nanopore_like(20_000, rng), 2, True)
    yield ol.stage1_encode((rng.integers(0, 1 << 20, 100_000)).astype(np.uint32), 4, False)
    x = np.cumsum(rng.integers(-3, 4, 900_000)).astype(np.int16)
    yield x.tobytes()


def test_restated_zstd_decoder_against_libzstd():
    rng = np.random.default_rng(11)
    n_frames = 0
    for data in _zstd_inputs(rng):
        for level in (1, 2, 3, 4, 5, 6, 9, 12, 16, 19, 22):
            if len(data) > 300_000 and level > 12:
                continue
            frame = ol.zstd_compress_lib(data, level)
            assert ol.frame_content_size(frame) == len(data)
            got = ol.zstd_decompress_port(frame, len(data))
            assert got == data, (len(data), level)
            n_frames += 1
            if len(frame) > 12:  # truncated and corrupted frames must be rejected or decode without crashing
                assert ol.zstd_decompress_port(frame[:-1], len(data)) is None
                assert ol.zstd_decompress_port(frame, len(data) - 1) is None
    assert n_frames > 100


def test_zstd_compress_bound():
    for n in (0, 1, 100, 131071, 131072, 131073, 10_000_000):
        want = n + (n >> 8) + (((128 << 10) - n) >> 11 if n < (128 << 10) else 0)
        assert port.lib.vbzo_zstd_compress_bound(n) == want
    assert ol.frame_content_size(b"\x28\xb5\x2f\xfd\x20\x00\x01\x00\x00") == 0
    assert ol.frame_content_size(b"\x28\xb5\x2f") == 2**64 - 2


def test_sized_header_quirks():
    o = Options(True, 2, 0, 0)
    arr = np.arange(10, dtype=np.int16)
    r, out = port.compress(arr, o, sized=True)
    assert struct.unpack("<I", out[:4])[0] == 20
    assert port.decompress(out[:3], o, 20)[0] == ol.ERR_INPUT_SIZE
    assert port.decompress(out, o, 19)[0] == ol.ERR_DEST_SIZE
    assert port.decompress(out, Options(True, 3, 0, 0), 20)[0] == ol.ERR_INTEGER_SIZE
    assert port.decompress(out, Options(True, 2, 0, 5), 20)[0] == ol.ERR_VERSION
    # odd byte count with integer_size 2: INPUT_SIZE_ERROR + 4 == 2
    r, _ = port.compress(np.zeros(3, np.uint8), o, sized=True, cap=64)
    assert r == 2
