#!/usr/bin/env python3
"""Emulator fuzz (no GPU needed; tests/cpusim build): the device-resident decode entry point with a tiny intermediate budget, so that every
call runs several waves with the look-ahead alternating between its two sets of buffers, ONE context for the whole run; chunks written by
the reference and by this library, reads that repeat short traces (a sequence every few bytes) among them.
usage: fuzz_device.py SEED SECONDS"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
os.environ["VBZ_B200_MID_MIB"] = "1"
import numpy as np, product, oracle_lib as ol, data_gen
from oracle_lib import Options
capi, batch, lib = product.get("cpusim")
ref = ol.ref() or ol.port()
rng = np.random.default_rng(int(sys.argv[1]))
t0 = time.time(); it = 0; bad = 0
while time.time() - t0 < float(sys.argv[2]):
    ver = int(rng.integers(0, 2))
    o = Options(True, 2, 1, ver)
    reads = []
    for _ in range(int(rng.integers(3, 14))):
        n = int(np.exp(rng.uniform(np.log(20), np.log(250_000))))
        k = rng.integers(0, 4)
        if k == 0: a = data_gen.nanopore_like(n, rng)
        elif k == 1: a = np.tile(data_gen.nanopore_like(int(rng.integers(5, 400)), rng), n // 5 + 1)[:n]
        elif k == 2: a = np.zeros(n, np.int16)
        else: a = rng.integers(-300, 300, n).astype(np.int16)
        reads.append(a)
    own = rng.random() < 0.5
    blobs = [b for _, b in batch.compress(reads, o, sized=True)] if own else [ref.compress(r, o, sized=True)[1] for r in reads]
    a16 = lambda v: (v + 15) & ~15
    so = np.cumsum([0] + [a16(len(b)) for b in blobs])[:-1].astype(np.uint64)
    do = np.cumsum([0] + [a16(r.nbytes) for r in reads])[:-1].astype(np.uint64)
    src = np.zeros(int(sum(a16(len(b)) for b in blobs)) + 16, np.uint8)
    dst = np.full(int(sum(a16(r.nbytes) for r in reads)) + 16, 0xA5, np.uint8)
    for b, off in zip(blobs, so):
        src[int(off):int(off) + len(b)] = np.frombuffer(b, np.uint8)
    res = np.zeros(len(reads), np.uint32)
    batch.decompress_device(src.ctypes.data, so, np.array([len(b) for b in blobs], np.uint32), dst.ctypes.data, do, np.array([r.nbytes for r in reads], np.uint32),
                            res.ctypes.data, o, True, 0)
    batch.synchronize()
    for i, (r, off) in enumerate(zip(reads, do)):
        if res[i] != r.nbytes or dst[int(off):int(off) + r.nbytes].tobytes() != r.tobytes():
            bad += 1; print("MISMATCH", "own" if own else "ref", i, r.size, hex(res[i]), [x.size for x in reads])
    it += 1
print("iterations", it, "bad", bad)
