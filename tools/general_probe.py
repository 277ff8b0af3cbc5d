"""Which frames of a host batch does the walker / the sequence kernel leave to the general kernel?  (GPU box or emulator)
python tools/general_probe.py [gpu|cpusim] [reads]"""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import data_gen  # noqa: E402
import product  # noqa: E402
from oracle_lib import Options  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "cpusim"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
capi, batch, lib = product.get(kind)
rng = np.random.default_rng(7)
lens = np.floor(np.exp(rng.uniform(np.log(1e4), np.log(1e6 if kind == "gpu" else 2e5), n))).astype(int)
reads = [data_gen.nanopore_like(int(k), rng) for k in lens]
o = Options(True, 2, 1, 1)
res = batch.compress(reads, o, sized=True)
stats = (ctypes.c_uint32 * 4)()
for group in (list(range(n)), ) + tuple([i] for i in range(n)):
    chunks = [res[i][1] for i in group]
    back = batch.decompress(chunks, [reads[i].nbytes for i in group], o, sized=True)
    ok = all(c == reads[i].nbytes and b == reads[i].tobytes() for i, (c, b) in zip(group, back))
    batch.lib.vbz_b200_decode_stats(batch.ctx, stats)
    if len(group) > 1 or stats[1]:
        print("reads", group if len(group) < 4 else "%d reads" % len(group), "samples", [int(lens[i]) for i in group][:4], "ok", ok,
              "units %d general %d seqframes %d records %d" % tuple(stats))
