"""Synthetic inputs shared by tests and bench (numpy, CPU).  The 'nanopore-like' model is the one
BASELINE.md section 3 / SURVEY.md 8(d) specify: level ~ N(400, 90) held for Geometric(0.1) samples,
sample = clip(level + N(0, 9), -4096, 4095)."""
import numpy as np


def nanopore_like(n, rng):
    if n == 0:
        return np.zeros(0, np.int16)
    change = rng.random(n) < 0.1
    change[0] = True
    seg = np.cumsum(change) - 1
    levels = rng.normal(400.0, 90.0, int(seg[-1]) + 1)
    x = levels[seg] + rng.normal(0.0, 9.0, n)
    return np.clip(np.rint(x), -4096, 4095).astype(np.int16)


def read_lengths(n_reads, rng, lo=1e4, hi=1e6):
    return np.floor(np.exp(rng.uniform(np.log(lo), np.log(hi), n_reads))).astype(np.int64)


def uniform_full_range(n, dtype, rng):
    info = np.iinfo(dtype)
    return rng.integers(info.min, info.max, n, dtype=dtype, endpoint=True)
