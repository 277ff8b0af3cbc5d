"""Per-call device time of encode / decode in a loop (device-resident batch), next to the host time of the call."""
import ctypes, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from vbz_compression_b200 import Batch, load
lib = load(); batch = Batch(0, lib=lib)
dev = torch.device("cuda")
wl = sys.argv[1] if len(sys.argv) > 1 else "full"
nreads = int(sys.argv[2]) if len(sys.argv) > 2 else 40000
opts = bench.options_for(wl)
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
K = 6
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * K + 1)]
host = []
torch.cuda.synchronize()
ev[0].record()
for k in range(K):
    t0 = time.perf_counter(); enc(); t1 = time.perf_counter(); ev[2 * k + 1].record()
    dec(); t2 = time.perf_counter(); ev[2 * k + 2].record()
    host.append((t1 - t0, t2 - t1))
torch.cuda.synchronize()
for k in range(K):
    print("step %d: encode dev %.2f ms (host call %.2f ms)  decode dev %.2f ms (host call %.2f ms)" % (k, ev[2 * k].elapsed_time(ev[2 * k + 1]), host[k][0] * 1e3, ev[2 * k + 1].elapsed_time(ev[2 * k + 2]), host[k][1] * 1e3))
