"""-m gpu: the device-resident batch C-ABI (vbz_b200_compress_device / vbz_b200_decompress_device) on a
real B200, torch used only for device memory.  Parity with the oracle on oracle-sized batches; at
benchmark-like sizes, size-independent properties (poisoned-output round trip, sum of result sizes)."""
import ctypes

import numpy as np
import pytest

import data_gen
import oracle_lib as ol
from oracle_lib import Options

pytestmark = pytest.mark.gpu


def _setup():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import product
    _, batch, lib = product.get("gpu")
    return torch, batch, lib


def _layout(lib, sizes, opts):
    a16 = lambda v: (int(v) + 15) & ~15
    src_off, dst_off, caps, so, do = [], [], [], 0, 0
    for s in sizes:
        src_off.append(so); so += a16(s)
        c = int(lib.vbz_b200_compress_bound(int(s), ctypes.byref(opts), 1))
        dst_off.append(do); caps.append(c); do += c
    return np.array(src_off, np.uint64), np.array(dst_off, np.uint64), np.array(caps, np.uint32), so, do


@pytest.mark.parametrize("dtype,opts", [
    (np.int16, Options(True, 2, 0, 0)), (np.int16, Options(False, 2, 0, 0)), (np.int32, Options(True, 4, 0, 0)),
    (np.uint32, Options(False, 4, 0, 0)), (np.int8, Options(True, 1, 0, 0)), (np.int8, Options(False, 1, 0, 0)),
    (np.int8, Options(True, 1, 0, 1)), (np.int8, Options(False, 1, 0, 1))])
def test_device_batch_matches_oracle(dtype, opts):
    torch, batch, lib = _setup()
    port = ol.port()
    rng = np.random.default_rng(21)
    lens = [0, 1, 5, 4096, 4097, 33333, 100_000, 8, 65536, 12, 250_001]
    if dtype == np.int16:
        reads = [data_gen.nanopore_like(n, rng) if i % 2 == 0 else data_gen.uniform_full_range(n, np.int16, rng) for i, n in enumerate(lens)]
    elif dtype == np.uint32:
        reads = [rng.integers(0, 1 << 20, n).astype(np.uint32) for n in lens]
    else:
        reads = [data_gen.uniform_full_range(n, dtype, rng) if i % 2 else rng.integers(-5, 6, n).astype(dtype) for i, n in enumerate(lens)]
    sizes = np.array([r.nbytes for r in reads], np.uint32)
    src_off, dst_off, caps, tot_s, tot_d = _layout(lib, sizes, opts)
    h_src = np.zeros(tot_s + 16, np.uint8)
    for r, o in zip(reads, src_off):
        h_src[int(o):int(o) + r.nbytes] = r.view(np.uint8)
    d_src = torch.from_numpy(h_src).cuda()
    d_cmp = torch.zeros(tot_d + 16, dtype=torch.uint8, device="cuda")
    d_back = torch.full((tot_s + 16,), 0xA5, dtype=torch.uint8, device="cuda")
    d_rc = torch.zeros(len(reads), dtype=torch.int32, device="cuda")
    d_rd = torch.zeros(len(reads), dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    batch.compress_device(d_src.data_ptr(), src_off, sizes, d_cmp.data_ptr(), dst_off, caps, d_rc.data_ptr(), opts, True, st)
    torch.cuda.synchronize()
    rc = d_rc.cpu().numpy().view(np.uint32)
    h_cmp = d_cmp.cpu().numpy()
    for i, r in enumerate(reads):
        want_code, want = port.compress(r, opts, sized=True)
        assert rc[i] == want_code, (i, rc[i], want_code)
        assert h_cmp[int(dst_off[i]):int(dst_off[i]) + int(rc[i])].tobytes() == want, i
    batch.decompress_device(d_cmp.data_ptr(), dst_off, rc, d_back.data_ptr(), src_off, sizes, d_rd.data_ptr(), opts, True, st)
    torch.cuda.synchronize()
    rd = d_rd.cpu().numpy().view(np.uint32)
    h_back = d_back.cpu().numpy()
    for i, r in enumerate(reads):
        assert rd[i] == r.nbytes
        assert h_back[int(src_off[i]):int(src_off[i]) + r.nbytes].tobytes() == r.tobytes(), i


def test_large_batch_properties():
    """~2 GB raw, 5000 reads of 10k-1M samples: every result is a size, the sizes add up to
    keys + data bytes counted independently with torch, and a poisoned output buffer comes back
    equal to the input everywhere except the inter-read padding."""
    torch, batch, lib = _setup()
    opts = Options(True, 2, 0, 0)
    rng = np.random.default_rng(20261017)
    lens = np.floor(np.exp(rng.uniform(np.log(1e4), np.log(1e6), 5000))).astype(np.int64)
    sizes = (lens * 2).astype(np.uint32)
    src_off, dst_off, caps, tot_s, tot_d = _layout(lib, sizes, opts)
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    x = torch.cumsum(torch.randint(-40, 41, (tot_s // 2,), device="cuda", generator=g, dtype=torch.int32), 0)
    d_src = ((x % 3000) + 200).to(torch.int16).view(torch.uint8)
    d_src[::977].bitwise_xor_(0x55)  # sprinkle large jumps -> two-byte values
    d_cmp = torch.empty(tot_d, dtype=torch.uint8, device="cuda")
    d_rc = torch.zeros(len(lens), dtype=torch.int32, device="cuda")
    d_rd = torch.zeros(len(lens), dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    batch.compress_device(d_src.data_ptr(), src_off, sizes, d_cmp.data_ptr(), dst_off, caps, d_rc.data_ptr(), opts, True, st)
    torch.cuda.synchronize()
    rc = d_rc.cpu().numpy().view(np.uint32)
    assert (rc < 0xFFFFFFF9).all()
    # independent count for a few reads: 4 + ceil(n/4) + n + #(zigzag(delta) > 255)
    s16 = d_src.view(torch.int16)
    for i in (0, 17, 4999):
        o, n = int(src_off[i]) // 2, int(lens[i])
        v = s16[o:o + n].to(torch.int32)
        d = torch.diff(v, prepend=torch.zeros(1, dtype=torch.int32, device="cuda"))
        d = ((d + 32768) % 65536) - 32768
        zz = (d * 2) ^ (d >> 31)
        assert int(rc[i]) == 4 + (n + 3) // 4 + n + int((zz > 255).sum().item())
    d_back = d_src.clone(); d_back.add_(1)
    batch.decompress_device(d_cmp.data_ptr(), dst_off, rc, d_back.data_ptr(), src_off, sizes, d_rd.data_ptr(), opts, True, st)
    torch.cuda.synchronize()
    assert (d_rd.cpu().numpy().view(np.uint32) == sizes).all()
    mism = int((d_src != d_back).sum().item())
    assert mism == tot_s - int(sizes.astype(np.int64).sum())


@pytest.mark.parametrize("level", [0, 1])
def test_torch_tensors_zero_copy(level):
    """reads that already live in CUDA tensors are compressed and decompressed where they are."""
    torch, batch, _ = _setup()
    from vbz_compression_b200 import CompressionOptions
    port = ol.port()
    rng = np.random.default_rng(21 + level)
    arrs = [data_gen.nanopore_like(int(n), rng) for n in (70_001, 8, 0, 12_345, 300_000)]
    reads = [torch.from_numpy(a.copy()).cuda() for a in arrs]
    opts = CompressionOptions(True, 2, level, 1)
    out, offs, sizes = batch.compress_tensors(reads, opts, sized=True)
    torch.cuda.synchronize()
    sz = sizes.cpu().numpy().view(np.uint32)
    assert (sz < 0xFFFFFFF9).all()
    o = Options(True, 2, level, 1)
    chunks = []
    for a, off, z in zip(arrs, offs, sz):
        chunk = out[int(off):int(off) + int(z)]
        if level == 0:
            rp, bp = port.compress(a, o, sized=True)
            assert int(z) == rp and chunk.cpu().numpy().tobytes() == bp
        chunks.append(chunk)
    back, res = batch.decompress_tensors(chunks, torch.int16, [a.size for a in arrs], opts, sized=True)
    torch.cuda.synchronize()
    assert (res.cpu().numpy().view(np.uint32) == np.array([a.nbytes for a in arrs], np.uint32)).all()
    for a, b in zip(arrs, back):
        assert np.array_equal(b.cpu().numpy(), a)


@pytest.mark.parametrize("level", [0, 1])
def test_chunks_of_non_x86_writers_decode_where_they_lie(level):
    """The reference built without SSSE3 (ARM, generic) writes 3-byte codes for int16 deltas beyond 16 bits: valid chunks whose stage-1
    payload exceeds the x86 encoder's keys + 2N.  The device-resident decoder cannot read frame headers on the host, so its
    intermediate slots must be sized by the format's bound (keys + 4N), not by what this library's own encoder would have written.
    What comes out is what the x86 reference makes of such a chunk (its 16-bit arithmetic wraps where the writer's 32-bit did not:
    full-range deltas do not survive the trip between the two builds of the reference either)."""
    torch, batch, _ = _setup()
    from vbz_compression_b200 import CompressionOptions
    gen = ol.ref(generic=True)
    if gen is None:
        pytest.skip("oracle/_ref/libvbz_ref_generic.so is not built")
    rng = np.random.default_rng(77 + level)
    arrs = [data_gen.uniform_full_range(n, np.int16, rng) for n in (9000, 33, 70_000)] + [data_gen.nanopore_like(20_000, rng)]
    o = Options(True, 2, level, 0)
    blobs = [gen.compress(a, o, sized=True)[1] for a in arrs]
    x86 = ol.ref()
    assert blobs[0] != x86.compress(arrs[0], o, sized=True)[1]  # (the two builds do write different streams for full-range deltas)
    want = [x86.decompress(b, o, a.nbytes, sized=True) for a, b in zip(arrs, blobs)]
    assert all(code == a.nbytes for (code, _), a in zip(want, arrs)) and want[3][1] == arrs[3].tobytes()
    chunks = [torch.from_numpy(np.frombuffer(b, np.uint8).copy()).cuda() for b in blobs]
    back, res = batch.decompress_tensors(chunks, torch.int16, [a.size for a in arrs], CompressionOptions(True, 2, level, 0), sized=True)
    torch.cuda.synchronize()
    assert (res.cpu().numpy().view(np.uint32) == np.array([a.nbytes for a in arrs], np.uint32)).all()
    for (_, w), b in zip(want, back):
        assert b.cpu().numpy().tobytes() == w


@pytest.mark.parametrize("lookahead", ["0", "1"])
def test_device_batches_in_waves(lookahead, monkeypatch):
    """A device-resident batch larger than the intermediate budget runs as waves; on the decode side the walker and the sequence kernel of
    wave w + 1 run on a second stream beside the kernels that write wave w's output (csrc/host/batch.cpp: decompress_device_impl).
    Results, the order with respect to the caller's stream and the per-stage timers must not depend on that."""
    torch, batch, lib = _setup()
    import ctypes
    from vbz_compression_b200 import CompressionOptions
    monkeypatch.setenv("VBZ_B200_MID_MIB", "1")  # a wave per one or two reads
    monkeypatch.setenv("VBZ_B200_DECODE_LOOKAHEAD", lookahead)
    rng = np.random.default_rng(33)
    arrs = [data_gen.nanopore_like(int(n), rng) for n in (260_000, 310_000, 5, 150_000, 401_000, 90_000, 222_222, 0, 333_333, 120_000, 64)]
    reads = [torch.from_numpy(a.copy()).cuda() for a in arrs]
    opts = CompressionOptions(True, 2, 1, 1)
    assert lib.vbz_b200_set_stage_timing(batch.ctx, 1) == 0
    out, offs, sizes = batch.compress_tensors(reads, opts, sized=True)
    sz = sizes.cpu().numpy().view(np.uint32)     # (the copy is queued behind the call on the same stream)
    assert (sz < 0xFFFFFFF9).all()
    t1, t2 = ctypes.c_float(), ctypes.c_float()
    assert lib.vbz_b200_stage_times(batch.ctx, ctypes.byref(t1), ctypes.byref(t2)) == 0 and t1.value > 0 and t2.value > 0
    chunks = [out[int(off):int(off) + int(z)] for off, z in zip(offs, sz)]
    port = ol.port()
    o = Options(True, 2, 1, 1)
    for a, c in zip(arrs, chunks):               # every frame decodes with the oracle
        rd, back = port.decompress(c.cpu().numpy().tobytes(), o, a.nbytes, sized=True)
        assert rd == a.nbytes and back == a.tobytes()
    for _ in range(2):                           # twice: the second call finds both sets of buffers in place
        back, res = batch.decompress_tensors(chunks, torch.int16, [a.size for a in arrs], opts, sized=True)
        assert (res.cpu().numpy().view(np.uint32) == np.array([a.nbytes for a in arrs], np.uint32)).all()
        assert lib.vbz_b200_stage_times(batch.ctx, ctypes.byref(t1), ctypes.byref(t2)) == 0 and t1.value > 0 and t2.value > 0
        for a, b in zip(arrs, back):
            assert np.array_equal(b.cpu().numpy(), a)
    lib.vbz_b200_set_stage_timing(batch.ctx, 0)
