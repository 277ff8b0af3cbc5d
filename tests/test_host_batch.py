"""The host-pointer batch entry points (include/vbz_b200.h) run a batch as pipelined waves
(csrc/host/batch.cpp: run_host_batch).  Every read must come back exactly as the oracle's per-read call
returns it, whatever the wave size, the path (staged through pinned memory / direct DMA from page-locked
caller buffers) and the mix of good and rejected reads."""
import ctypes
import os

import numpy as np
import pytest

import data_gen
import oracle_lib as ol
import product
from oracle_lib import Options
from vbz_compression_b200 import CompressionOptions

port = ol.port()


@pytest.fixture(params=product.BACKENDS)
def backend(request):
    return product.get(request.param)


def _reads(rng, n_reads, lo, hi):
    return [data_gen.nanopore_like(int(n), rng) for n in rng.integers(lo, hi, n_reads)]


@pytest.mark.parametrize("level", [0, 1])
@pytest.mark.parametrize("wave_mib", ["1", "0"])
def test_waves_match_oracle(backend, level, wave_mib, monkeypatch):
    _, batch, _ = backend
    if wave_mib != "0":
        monkeypatch.setenv("VBZ_B200_WAVE_MIB", wave_mib)
    else:
        monkeypatch.delenv("VBZ_B200_WAVE_MIB", raising=False)
    rng = np.random.default_rng(11 + level)
    reads = _reads(rng, 40, 1, 120_000)
    reads[5] = np.zeros(0, np.int16)
    reads[17] = reads[17].view(np.uint8)[:-1].copy()  # odd byte count: VBZ_INPUT_SIZE_ERROR for this read only
    o = Options(True, 2, level, 1)
    opts = CompressionOptions(True, 2, level, 1)
    res = batch.compress(reads, opts, sized=True)
    blobs = []
    for arr, (code, blob) in zip(reads, res):
        cap = port.max_compressed_size(arr.nbytes, o)
        if ol.is_error(cap):  # api.Batch hands such a read a 4-byte slot; the sized entry point adds 4 to the error (vbz.cpp:329)
            rp, _ = port.compress(arr, o, sized=True, cap=4)
            assert code == rp == (ol.ERR_INPUT_SIZE + 4) & 0xFFFFFFFF
            blobs.append(b"\x00" * 8)
            continue
        rp, bp = port.compress(arr, o, sized=True)
        assert not ol.is_error(code) and not ol.is_error(rp)
        if level == 0:
            assert code == rp and blob == bp
        else:  # zstd bytes are the encoder's own; the oracle must decode them
            r2, back = port.decompress(blob, o, arr.nbytes, sized=True)
            assert r2 == arr.nbytes and back == arr.tobytes()
        blobs.append(blob)
    blobs[9] = blobs[9][:len(blobs[9]) // 2]  # a truncated chunk: fails alone
    back = batch.decompress(blobs, [a.nbytes for a in reads], opts, sized=True)
    for i, (arr, (code, blob)) in enumerate(zip(reads, back)):
        rp, bp = port.decompress(blobs[i], o, arr.nbytes, sized=True)
        assert code == rp, i
        if not ol.is_error(rp):
            assert blob == bp


@pytest.mark.parametrize("upload", ["kernel", "dma"])
@pytest.mark.parametrize("level", [0, 1])
def test_transfer_kernels_any_alignment(level, upload, monkeypatch):
    """Page-locked caller buffers are read and written in place by the transfer kernels (csrc/xfer_kernels.cuh), whatever
    their byte alignment.  Runs on the emulator, which shares one address space with its "device"."""
    _, batch, lib = product.get("cpusim")
    monkeypatch.setenv("VBZ_CPUSIM_PINNED", "1")
    monkeypatch.setenv("VBZ_B200_UPLOAD", upload)
    monkeypatch.setenv("VBZ_B200_WAVE_MIB", "1")
    rng = np.random.default_rng(77 + level)
    lens = [int(x) for x in rng.integers(1, 60_000, 24)] + [0, 7, 8, 9, 40_000]
    o = Options(True, 2, level, 1)
    opts = CompressionOptions(True, 2, level, 1)
    n = len(lens)
    reads = [data_gen.nanopore_like(x, rng) for x in lens]
    # sources back to back from an odd address (two runs: a gap after read 10); destinations at odd, spaced addresses
    gaps = [5 if i == 11 else 0 for i in range(n)]
    offs = np.cumsum([1] + [2 * lens[i] + gaps[i + 1] if i + 1 < n else 0 for i in range(n)])[:n]
    src = np.zeros(int(offs[-1]) + 2 * lens[-1] + 64, np.uint8)
    for a, off in zip(reads, offs):
        src[int(off):int(off) + a.nbytes] = a.view(np.uint8)
    caps = [int(lib.vbz_max_compressed_size(2 * x, ctypes.byref(opts))) for x in lens]
    coffs = np.cumsum([3] + [c + 3 for c in caps])[:n]
    comp = np.full(int(coffs[-1]) + caps[-1] + 64, 0xEE, np.uint8)
    back = np.full(src.size, 0xDD, np.uint8)
    vp, u32 = ctypes.c_void_p, ctypes.c_uint32
    srcp = (vp * n)(*[src.ctypes.data + int(x) for x in offs])
    cp = (vp * n)(*[comp.ctypes.data + int(x) for x in coffs])
    bp = (vp * n)(*[back.ctypes.data + int(x) for x in offs])
    ssz = (u32 * n)(*[2 * x for x in lens])
    ccap = (u32 * n)(*caps)
    res_c, res_d = (u32 * n)(), (u32 * n)()
    l0 = batch.kernel_launches
    assert lib.vbz_b200_compress_batch(batch.ctx, srcp, ssz, cp, ccap, res_c, n, ctypes.byref(opts), 1) == 0, lib.vbz_b200_last_error()
    assert lib.vbz_b200_decompress_batch(batch.ctx, cp, res_c, bp, ssz, res_d, n, ctypes.byref(opts), 1) == 0, lib.vbz_b200_last_error()
    assert batch.kernel_launches > l0
    for i, a in enumerate(reads):
        assert res_d[i] == a.nbytes
        blob = comp[int(coffs[i]):int(coffs[i]) + res_c[i]].tobytes()
        if level == 0:
            rp, bpo = port.compress(a, o, sized=True)
            assert res_c[i] == rp and blob == bpo
        else:
            r2, dec = port.decompress(blob, o, a.nbytes, sized=True)
            assert r2 == a.nbytes and dec == a.tobytes()
        assert back[int(offs[i]):int(offs[i]) + a.nbytes].tobytes() == a.tobytes()
        # nothing outside a read's own bytes is written
        end = int(coffs[i]) + res_c[i]
        assert (comp[end:end + 3] == 0xEE).all() and (comp[int(coffs[i]) - 3:int(coffs[i])] == 0xEE).all()
    assert back[0] == 0xDD and (back[int(offs[-1]) + 2 * lens[-1]:] == 0xDD).all()


@pytest.mark.gpu
@pytest.mark.parametrize("level", [0, 1])
def test_direct_dma_from_pinned_buffers(level):
    """Page-locked caller buffers are read and written in place by the device (transfer kernels, or run-by-run DMA
    for host-contiguous inputs), no staging copy."""
    import torch
    _, batch, lib = product.get("gpu")
    rng = np.random.default_rng(5)
    lens = rng.integers(40_000, 400_000, 64)
    opts = CompressionOptions(True, 2, level, 0)
    tot = int(lens.sum()) * 2
    host_in = torch.empty(tot, dtype=torch.uint8).pin_memory()
    arr = data_gen.nanopore_like(int(lens.sum()), rng)
    host_in.numpy()[:] = arr.view(np.uint8)
    caps = [int(lib.vbz_max_compressed_size(int(n) * 2, ctypes.byref(opts))) for n in lens]
    host_c = torch.zeros(int(sum(caps)), dtype=torch.uint8).pin_memory()
    host_back = torch.zeros(tot, dtype=torch.uint8).pin_memory()
    n = len(lens)
    vp, u32 = ctypes.c_void_p, ctypes.c_uint32
    offs = np.concatenate([[0], np.cumsum(lens * 2)[:-1]])
    coffs = np.concatenate([[0], np.cumsum(caps)[:-1]])
    srcp = (vp * n)(*[host_in.data_ptr() + int(o) for o in offs])
    cp = (vp * n)(*[host_c.data_ptr() + int(o) for o in coffs])
    bp = (vp * n)(*[host_back.data_ptr() + int(o) for o in offs])
    ssz = (u32 * n)(*[int(x) * 2 for x in lens])
    ccap = (u32 * n)(*caps)
    res_c, res_d = (u32 * n)(), (u32 * n)()
    for path in ("kernel", "dma", "staged"):
        os.environ["VBZ_B200_HOST_PATH"] = "staged" if path == "staged" else "auto"
        os.environ["VBZ_B200_UPLOAD"] = "dma" if path == "dma" else "kernel"
        os.environ["VBZ_B200_WAVE_MIB"] = "8"
        try:
            host_c.zero_(); host_back.zero_()
            assert lib.vbz_b200_compress_batch(batch.ctx, srcp, ssz, cp, ccap, res_c, n, ctypes.byref(opts), 1) == 0, lib.vbz_b200_last_error()
            assert lib.vbz_b200_decompress_batch(batch.ctx, cp, res_c, bp, ssz, res_d, n, ctypes.byref(opts), 1) == 0, lib.vbz_b200_last_error()
        finally:
            del os.environ["VBZ_B200_HOST_PATH"], os.environ["VBZ_B200_WAVE_MIB"], os.environ["VBZ_B200_UPLOAD"]
        assert list(res_d) == [int(x) * 2 for x in lens]
        assert torch.equal(host_in, host_back), path
        if level == 0:
            o = Options(True, 2, 0, 0)
            for i in (0, n // 2, n - 1):
                a = arr[int(offs[i]) // 2: int(offs[i]) // 2 + int(lens[i])]
                rp, bpo = port.compress(a, o, sized=True)
                got = host_c.numpy()[int(coffs[i]): int(coffs[i]) + res_c[i]].tobytes()
                assert res_c[i] == rp and got == bpo


@pytest.mark.gpu
@pytest.mark.parametrize("level", [0, 1])
def test_pinned_and_pageable_reads_in_one_wave(level):
    """Every read chooses its own way over the link: page-locked buffers are moved in place by the transfer kernels, pageable
    ones through the wave's pinned staging - in the same wave, next to rejected and empty reads."""
    import torch
    _, batch, lib = product.get("gpu")
    rng = np.random.default_rng(91 + level)
    lens = [int(x) for x in rng.integers(1_000, 300_000, 40)]
    lens[3] = 0
    arrs = [data_gen.nanopore_like(n, rng) for n in lens]
    opts = CompressionOptions(True, 2, level, 1)
    o = Options(True, 2, level, 1)
    n = len(arrs)
    keep = []
    srcp, cp, bp, ssz, ccap = [], [], [], [], []
    for i, a in enumerate(arrs):
        nbytes = a.nbytes - (1 if i == 7 else 0)  # read 7: odd byte count -> VBZ_INPUT_SIZE_ERROR for it alone
        cap = int(lib.vbz_max_compressed_size(a.nbytes, ctypes.byref(opts)))
        if i % 2 == 0:  # page-locked, deliberately misaligned by 2 bytes
            s = torch.empty(a.nbytes + 16, dtype=torch.uint8).pin_memory()
            s.numpy()[2:2 + a.nbytes] = a.view(np.uint8)
            c = torch.zeros(cap + 16, dtype=torch.uint8).pin_memory()
            b = torch.zeros(a.nbytes + 16, dtype=torch.uint8).pin_memory()
            keep += [s, c, b]
            srcp.append(s.data_ptr() + 2); cp.append(c.data_ptr() + 6); bp.append(b.data_ptr() + 2)
        else:  # pageable numpy memory
            s = np.ascontiguousarray(a).view(np.uint8).copy() if a.size else np.zeros(1, np.uint8)
            c = np.zeros(cap + 16, np.uint8)
            b = np.zeros(a.nbytes + 16, np.uint8)
            keep += [s, c, b]
            srcp.append(s.ctypes.data); cp.append(c.ctypes.data); bp.append(b.ctypes.data)
        ssz.append(nbytes); ccap.append(cap)
    vp, u32 = ctypes.c_void_p, ctypes.c_uint32
    A = lambda t, v: (t * n)(*v)
    res_c, res_d = (u32 * n)(), (u32 * n)()
    os.environ["VBZ_B200_WAVE_MIB"] = "4"
    try:
        assert lib.vbz_b200_compress_batch(batch.ctx, A(vp, srcp), A(u32, ssz), A(vp, cp), A(u32, ccap), res_c, n, ctypes.byref(opts), 1) == 0
        dsz = [a.nbytes for a in arrs]
        assert lib.vbz_b200_decompress_batch(batch.ctx, A(vp, cp), res_c, A(vp, bp), A(u32, dsz), res_d, n, ctypes.byref(opts), 1) == 0
    finally:
        del os.environ["VBZ_B200_WAVE_MIB"]
    for i, a in enumerate(arrs):
        if i == 7:
            assert res_c[i] == (ol.ERR_INPUT_SIZE + 4) & 0xFFFFFFFF
            continue
        rp, bpo = port.compress(a, o, sized=True)
        got = bytes((ctypes.c_uint8 * res_c[i]).from_address(cp[i]))
        if level == 0:
            assert res_c[i] == rp and got == bpo, i
        else:
            r2, dec = port.decompress(got, o, a.nbytes, sized=True)
            assert r2 == a.nbytes and dec == a.tobytes(), i
        assert res_d[i] == a.nbytes
        assert bytes((ctypes.c_uint8 * a.nbytes).from_address(bp[i])) == a.tobytes(), i


def test_frames_that_claim_gigabytes_fail_alone(backend):
    """A chunk whose zstd frame header claims 2 GiB of content must cost the batch nothing: its intermediate slot is bounded by what
    stage 1 could consume for the caller's output size, the chunk gets its own error, the reads around it decode (the reference fails
    only the offending read: vbz.cpp:222-236)."""
    _, batch, _ = backend
    rng = np.random.default_rng(5)
    o = Options(True, 2, 1, 0)
    opts = CompressionOptions(True, 2, 1, 0)
    good = [data_gen.nanopore_like(n, rng) for n in (30_000, 7, 120_000)]
    blobs = [port.compress(a, o, sized=True)[1] for a in good]
    # sized header (40 bytes of samples) + frame header: magic, single-segment with a 4-byte content size of 0x7FFFFFF0, one empty raw last block
    crafted = (40).to_bytes(4, "little") + bytes([0x28, 0xB5, 0x2F, 0xFD, 0xA0]) + (0x7FFFFFF0).to_bytes(4, "little") + bytes([0x01, 0x00, 0x00])
    chunks = [blobs[0]] + [crafted] * 200 + blobs[1:]
    sizes = [good[0].nbytes] + [40] * 200 + [a.nbytes for a in good[1:]]
    res = batch.decompress(chunks, sizes, opts, sized=True)
    assert res[0] == (good[0].nbytes, good[0].tobytes())
    assert res[201] == (good[1].nbytes, good[1].tobytes()) and res[202] == (good[2].nbytes, good[2].tobytes())
    want, _ = port.decompress(crafted, o, 40, sized=True)
    assert ol.is_error(want)
    for code, blob in res[1:201]:
        assert ol.is_error(code) and blob == b""
