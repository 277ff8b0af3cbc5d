#!/usr/bin/env python3
"""Compression ratios of the product's encoder next to the reference's, on the inputs VERDICT.md names: the fast5 fixture reads,
numpy.arange(1000) (pyvbz README), a 2,000-sample trace tiled x100 (the shape of vbz/perf/test_data_generator.h), the bench signal.
Runs on the emulator build (no GPU needed): python tools/ratio_probe.py [gpu]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import data_gen  # noqa: E402
import golden_data  # noqa: E402
import oracle_lib as ol  # noqa: E402
import product  # noqa: E402
from oracle_lib import Options  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "cpusim"
capi, batch, lib = product.get(kind)
ref = ol.ref() or ol.port()


def both(arrs, o, name):
    ours = batch.compress(arrs, o, sized=True)
    raw = sum(a.nbytes for a in arrs)
    a = sum(len(b) for _, b in ours)
    b = 0
    for arr, (code, blob) in zip(arrs, ours):
        rc, theirs = ref.compress(arr, o, sized=True)
        b += len(theirs)
        rd, back = ref.decompress(blob, o, arr.nbytes, sized=True)
        assert rd == arr.nbytes and back == arr.tobytes(), (name, "reference cannot decode our frame")
    back = batch.decompress([bl for _, bl in ours], [x.nbytes for x in arrs], o, sized=True)
    for arr, (code, blob) in zip(arrs, back):
        assert code == arr.nbytes and blob == arr.tobytes(), (name, "own decode")
    print("%-34s raw %9d  ours %8d (%.5f)  reference %8d (%.5f)  ours/ref %.3f" % (name, raw, a, a / raw, b, b / raw, a / max(b, 1)))
    return a, b


for level in (1, 3):
    o = Options(True, 2, level, 1)
    print("level", level)
    reads = [rd["samples"] for rd in golden_data.fast5_reads()]
    both(reads, o, "fast5 fixtures (10 reads)")
    both([np.arange(1000, dtype=np.int16)], o, "arange(1000, int16)")
    rng = np.random.default_rng(3)
    both([np.tile(data_gen.nanopore_like(2000, rng), 100)], o, "2000-sample trace x100")
    both([np.tile(reads[0][:5000], 30)], o, "fixture trace 5000 x30")
    both([data_gen.nanopore_like(n, rng) for n in (20000, 150000, 400000)], o, "bench signal (3 reads)")
    both([np.zeros(100000, np.int16), np.full(30000, 7, np.int16)], o, "constant")
    both([rng.integers(-32768, 32767, 50000).astype(np.int16)], o, "uniform int16")
o = Options(False, 4, 3, 0)
both([np.random.default_rng(1).integers(0, 1 << 20, 100000).astype(np.uint32)], o, "uint32 U[0,2^20) zstd 3")
both([np.sort(np.random.default_rng(1).integers(0, 1 << 28, 100000)).astype(np.uint32)], o, "uint32 sorted ids zstd 3")
