"""End-to-end probe (GPU box): host-pointer batches through vbz_b200_{compress,decompress}_batch, timed per direction,
next to what the PCIe link itself delivers for the same bytes.  python tools/e2e_probe.py [gib] [workload ...]"""
import ctypes
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from vbz_compression_b200 import Batch, load  # noqa: E402


def pcie_probe():
    n = 1 << 30
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for name, fn in (("h2d", lambda: d.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(d, non_blocking=True))):
        fn(); torch.cuda.synchronize()
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print("pcie %s %.1f GB/s" % (name, n / dt / 1e9))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.cuda.stream(s1):
        d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2):
        h2.copy_(d2, non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("pcie both directions at once: %.1f GB/s each" % (n / dt / 1e9))


def main():
    gib = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
    workloads = sys.argv[2:] or ["transform", "full"]
    pcie_probe()
    lib = load()
    batch = Batch(0, lib=lib)
    lens = bench.read_lengths(20000, bench.SEED)
    src_size = (lens * 2).astype(np.uint32)
    take, tot = 0, 0
    while tot + int(src_size[take]) <= int(gib * 2**30):
        tot += int(src_size[take]); take += 1
    d = torch.empty(tot // 2, dtype=torch.int16, device="cuda")
    bench.nanopore_like_torch(d, bench.SEED)
    pageable = os.environ.get("PAGEABLE") is not None  # plain malloc'd caller buffers instead of page-locked ones
    host_in = torch.empty(tot, dtype=torch.uint8) if pageable else torch.empty(tot, dtype=torch.uint8).pin_memory()
    host_in.copy_(d.view(torch.uint8))
    del d
    offs = np.concatenate([[0], np.cumsum(src_size[:take].astype(np.int64))[:-1]])
    for wl in workloads:
        opts = bench.options_for(wl)
        caps = [int(lib.vbz_max_compressed_size(int(src_size[i]), ctypes.byref(opts))) for i in range(take)]
        coffs = np.concatenate([[0], np.cumsum(caps)[:-1]]).astype(np.int64)
        host_c = torch.zeros(int(sum(caps)), dtype=torch.uint8) if pageable else torch.empty(int(sum(caps)), dtype=torch.uint8).pin_memory()
        host_back = torch.zeros(tot, dtype=torch.uint8) if pageable else torch.empty(tot, dtype=torch.uint8).pin_memory()
        vp, u32 = ctypes.c_void_p, ctypes.c_uint32
        srcp = (vp * take)(*[host_in.data_ptr() + int(o) for o in offs])
        cp = (vp * take)(*[host_c.data_ptr() + int(o) for o in coffs])
        bp = (vp * take)(*[host_back.data_ptr() + int(o) for o in offs])
        ssz = (u32 * take)(*[int(src_size[i]) for i in range(take)])
        ccap = (u32 * take)(*caps)
        res_c, res_d = (u32 * take)(), (u32 * take)()

        def enc():
            assert lib.vbz_b200_compress_batch(batch.ctx, srcp, ssz, cp, ccap, res_c, take, ctypes.byref(opts), 1) == 0

        def dec():
            assert lib.vbz_b200_decompress_batch(batch.ctx, cp, res_c, bp, ssz, res_d, take, ctypes.byref(opts), 1) == 0

        for wave in os.environ.get("WAVES", "0").split(","):
            if wave != "0":
                os.environ["VBZ_B200_WAVE_MIB"] = wave
            else:
                os.environ.pop("VBZ_B200_WAVE_MIB", None)
            enc(); dec()
            assert torch.equal(host_in, host_back)
            comp = int(np.frombuffer(res_c, np.uint32).astype(np.int64).sum())
            te, td = [], []
            for _ in range(3):
                t0 = time.perf_counter(); enc(); t1 = time.perf_counter(); dec(); t2 = time.perf_counter()
                te.append(t1 - t0); td.append(t2 - t1)
            e, dd = min(te), min(td)
            print("%s wave=%s reads=%d raw=%.2f GiB comp=%.3f: encode %.1f ms (%.1f GB/s raw; link needs %.1f ms at 55 GB/s) decode %.1f ms (%.1f GB/s) combined %.1f GB/s"
                  % (wl, wave, take, tot / 2**30, comp / tot, e * 1e3, tot / e / 1e9, tot / 55e9 * 1e3, dd * 1e3, tot / dd / 1e9, 2 * tot / (e + dd) / 1e9))


if __name__ == "__main__":
    main()
