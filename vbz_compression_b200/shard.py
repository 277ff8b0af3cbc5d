"""Sharding a batch of reads over the GPUs of a box: every read is independent (the reference keeps no
state between calls, vbz/vbz.cpp), so rank k takes a contiguous range of the read index, chosen so that BYTES - not
counts - are balanced.  No data-path collective exists or is needed; outputs stay per read."""
import numpy as np


def balanced_ranges(sizes, world_size):
    """-> list of (first, last_exclusive) per rank: contiguous, covering, with byte sums as equal as whole reads allow."""
    sizes = np.asarray(sizes, dtype=np.int64)
    n = len(sizes)
    if world_size <= 0:
        raise ValueError("world_size must be positive")
    ends = np.cumsum(sizes)
    total = int(ends[-1]) if n else 0
    cuts = [0]
    for k in range(1, world_size):
        target = total * k / world_size
        # first index whose cumulative end passes the target; pick the closer of the two neighbouring cut points
        i = int(np.searchsorted(ends, target, side="left")) if n else 0
        if i < n:
            before = int(ends[i - 1]) if i > 0 else 0
            if abs(int(ends[i]) - target) < abs(before - target):
                i += 1
        cuts.append(min(max(i, cuts[-1]), n))
    cuts.append(n)
    return [(cuts[k], cuts[k + 1]) for k in range(world_size)]
