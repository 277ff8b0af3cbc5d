"""GPU debugging aid: run the bench's device-resident round trip at a given read count and report
where (which reads / byte positions) the round trip first differs."""
import ctypes, sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
from vbz_compression_b200 import Batch, load, CompressionOptions

def run(n_reads):
    lib = load(); batch = Batch(0, lib=lib)
    opts = CompressionOptions(True, 2, 0, 0)
    lens = bench.read_lengths(n_reads, bench.SEED)
    src_size = (lens * 2).astype(np.uint32)
    a16 = lambda v: (v + np.uint64(15)) & ~np.uint64(15)
    src_slot = a16(src_size.astype(np.uint64))
    src_off = np.concatenate([[0], np.cumsum(src_slot)[:-1]]).astype(np.uint64)
    caps = np.array([lib.vbz_b200_compress_bound(int(s), ctypes.byref(opts), 1) for s in src_size], np.uint64)
    dst_off = np.concatenate([[0], np.cumsum(caps)[:-1]]).astype(np.uint64)
    total_src, total_dst = int(src_slot.sum()), int(caps.sum())
    print("reads", n_reads, "total_src", total_src, "total_dst", total_dst, "dtypes", src_off.dtype, dst_off.dtype, src_slot.dtype)
    d_src = torch.empty(total_src, dtype=torch.uint8, device="cuda")
    bench.nanopore_like_torch(d_src.view(torch.int16), bench.SEED)
    d_cmp = torch.zeros(total_dst, dtype=torch.uint8, device="cuda")
    d_back = torch.zeros(total_src, dtype=torch.uint8, device="cuda")
    rc = torch.zeros(n_reads, dtype=torch.int32, device="cuda"); rd = torch.zeros(n_reads, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    batch.compress_device(d_src.data_ptr(), src_off, src_size, d_cmp.data_ptr(), dst_off, caps.astype(np.uint32), rc.data_ptr(), opts, True, st)
    torch.cuda.synchronize()
    res_c = rc.cpu().numpy().view(np.uint32)
    print("encode errors:", int((res_c >= 0xFFFFFFF9).sum()))
    batch.decompress_device(d_cmp.data_ptr(), dst_off, res_c, d_back.data_ptr(), src_off, src_size, rd.data_ptr(), opts, True, st)
    torch.cuda.synchronize()
    res_d = rd.cpu().numpy().view(np.uint32)
    print("decode size mismatches:", int((res_d != src_size).sum()))
    bad = []
    for i in range(n_reads):
        o, s = int(src_off[i]), int(src_size[i])
        if not torch.equal(d_src[o:o + s], d_back[o:o + s]):
            bad.append(i)
            if len(bad) <= 5:
                a = d_src[o:o + s].view(torch.int16); b = d_back[o:o + s].view(torch.int16)
                w = torch.nonzero(a != b).flatten()
                print(" bad read", i, "src_off", o, "samples", s // 2, "first diff sample", int(w[0]), "ndiff", int(w.numel()),
                      "a", a[int(w[0]):int(w[0]) + 4].tolist(), "b", b[int(w[0]):int(w[0]) + 4].tolist(), "dst_off", int(dst_off[i]), "csize", int(res_c[i]))
                # independent size
                v = a.to(torch.int32); d = torch.diff(v, prepend=torch.zeros(1, dtype=torch.int32, device="cuda")); d = ((d + 32768) % 65536) - 32768
                zz = (d * 2) ^ (d >> 31); n = s // 2
                print("   expected csize", 4 + (n + 3) // 4 + n + int((zz > 255).sum().item()))
    print("bad reads:", len(bad), "first", bad[:10], "last", bad[-3:])

if __name__ == "__main__":
    for n in [int(x) for x in sys.argv[1:]]:
        run(n)
        torch.cuda.empty_cache()
