import ctypes, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, data_gen, product
import oracle_lib as ol
from oracle_lib import Options
capi, batch, lib = product.get(sys.argv[1] if len(sys.argv) > 1 else "cpusim")
ref = ol.ref() or ol.port()
rng = np.random.default_rng(7)
n = 24
lens = np.floor(np.exp(rng.uniform(np.log(1e4), np.log(1e6 if len(sys.argv) > 1 and sys.argv[1] == "gpu" else 2e5), n))).astype(int)
reads = [data_gen.nanopore_like(int(k), rng) for k in lens]
o = Options(True, 2, 1, 1)
chunks = [ref.compress(a, o, sized=True)[1] for a in reads]
stats = (ctypes.c_uint32 * 4)()
back = batch.decompress(chunks, [a.nbytes for a in reads], o, sized=True)
assert all(c == a.nbytes and b == a.tobytes() for a, (c, b) in zip(reads, back))
batch.lib.vbz_b200_decode_stats(batch.ctx, stats)
tot = sum(a.nbytes for a in reads)
print("batch of %d reference-written chunks: units %d general %d seqframes %d records %d; raw bytes per record %.1f, content bytes per record %.1f" % ((n,) + tuple(stats) + (tot / max(stats[3], 1), 0.64 * tot / max(stats[3], 1))))
worst = 0
for i in range(n):
    back = batch.decompress([chunks[i]], [reads[i].nbytes], o, sized=True)
    batch.lib.vbz_b200_decode_stats(batch.ctx, stats)
    if stats[1]: print("  read", i, lens[i], "samples -> general kernel", tuple(stats))
