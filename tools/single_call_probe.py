"""Latency of the single-read C API (what libhdf5 calls once per chunk through the H5Z filter)."""
import ctypes, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import data_gen
from vbz_compression_b200 import CompressionOptions, load
lib = load()
rng = np.random.default_rng(1)
for level in (0, 1):
    o = CompressionOptions(True, 2, level, 1)
    for n in (10_000, 100_000, 1_000_000):
        a = data_gen.nanopore_like(n, rng)
        cap = lib.vbz_max_compressed_size(a.nbytes, ctypes.byref(o))
        out = np.empty(cap, np.uint8); back = np.empty(a.nbytes, np.uint8)
        for _ in range(3):
            r = lib.vbz_compress_sized(a.ctypes.data, a.nbytes, out.ctypes.data, cap, ctypes.byref(o))
            r2 = lib.vbz_decompress_sized(out.ctypes.data, r, back.ctypes.data, a.nbytes, ctypes.byref(o))
        assert r2 == a.nbytes and back.tobytes() == a.tobytes()
        K = 50
        t0 = time.perf_counter()
        for _ in range(K): lib.vbz_compress_sized(a.ctypes.data, a.nbytes, out.ctypes.data, cap, ctypes.byref(o))
        t1 = time.perf_counter()
        for _ in range(K): lib.vbz_decompress_sized(out.ctypes.data, r, back.ctypes.data, a.nbytes, ctypes.byref(o))
        t2 = time.perf_counter()
        print("level %d, %7d samples: compress %.0f us (%.2f GB/s)  decompress %.0f us (%.2f GB/s)" % (level, n, (t1 - t0) / K * 1e6, a.nbytes * K / (t1 - t0) / 1e9, (t2 - t1) / K * 1e6, a.nbytes * K / (t2 - t1) / 1e9))
