"""How long do the codec kernels take on a wave-sized, device-resident batch?  python tools/wave_probe.py"""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from vbz_compression_b200 import Batch, load
lib = load(); batch = Batch(0, lib=lib)
dev = torch.device("cuda")
for wl in ("transform", "full"):
    opts = bench.options_for(wl)
    for nreads in (72, 288, 2304):
        lens = bench.read_lengths(nreads, 1); src_size = (lens * 2).astype(np.uint32)
        a16 = lambda v: (v + 15) & ~np.uint64(15)
        slot = a16(src_size.astype(np.uint64)); src_off = np.concatenate([[0], np.cumsum(slot)[:-1]]).astype(np.uint64)
        caps = np.array([lib.vbz_b200_compress_bound(int(s), ctypes.byref(opts), 1) for s in src_size], np.uint64)
        dst_off = np.concatenate([[0], np.cumsum(caps)[:-1]]).astype(np.uint64)
        d_src = torch.empty(int(slot.sum()), dtype=torch.uint8, device=dev); bench.nanopore_like_torch(d_src.view(torch.int16), 1)
        d_cmp = torch.empty(int(caps.sum()), dtype=torch.uint8, device=dev); d_back = torch.empty_like(d_src)
        rc = torch.zeros(len(lens), dtype=torch.int32, device=dev); rd = torch.zeros_like(rc)
        st = torch.cuda.current_stream().cuda_stream
        caps32 = caps.astype(np.uint32); raw = int(src_size.astype(np.int64).sum())
        def enc(): batch.compress_device(d_src.data_ptr(), src_off, src_size, d_cmp.data_ptr(), dst_off, caps32, rc.data_ptr(), opts, True, st)
        enc(); torch.cuda.synchronize(); cs = rc.cpu().numpy().view(np.uint32).copy()
        def dec(): batch.decompress_device(d_cmp.data_ptr(), dst_off, cs, d_back.data_ptr(), src_off, src_size, rd.data_ptr(), opts, True, st)
        for f in (enc, dec):
            for _ in range(3): f()
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20): f()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            print("%s %d reads %.1f MB raw: %s %.3f ms  (%.0f GB/s raw)" % (wl, nreads, raw / 1e6, f.__name__, ms, raw / ms / 1e6))
