"""The multi-GPU path: reads sharded by index over ranks, bytes balanced, no data-path collective.
Run here with world_size 2 on the gloo backend; every rank drives the emulator build of the library."""
import os
import socket
import sys
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from vbz_compression_b200.shard import balanced_ranges  # noqa: E402


def test_balanced_ranges_cover_and_balance():
    rng = np.random.default_rng(3)
    sizes = np.floor(np.exp(rng.uniform(np.log(2e4), np.log(2e6), 1000))).astype(np.int64)
    for world in (1, 2, 4, 8):
        rs = balanced_ranges(sizes, world)
        assert rs[0][0] == 0 and rs[-1][1] == len(sizes)
        assert all(rs[k][1] == rs[k + 1][0] for k in range(world - 1))
        loads = [int(sizes[a:b].sum()) for a, b in rs]
        assert max(loads) - min(loads) <= 2 * int(sizes.max())
    assert balanced_ranges([], 2) == [(0, 0), (0, 0)]
    assert balanced_ranges([5], 4)[0][0] == 0 and balanced_ranges([5], 4)[-1][1] == 1


def _worker(rank, world, port, out_q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import data_gen
    import product
    from vbz_compression_b200 import CompressionOptions
    _, batch, _ = product.get("cpusim")
    rng = np.random.default_rng(99)  # the same batch on every rank; each takes its own range of it
    lens = rng.integers(1, 60_000, 24)
    reads = [data_gen.nanopore_like(int(n), rng) for n in lens]
    a, b = balanced_ranges([r.nbytes for r in reads], world)[rank]
    opts = CompressionOptions(True, 2, 1, 1)
    res = batch.compress(reads[a:b], opts, sized=True)
    back = batch.decompress([blob for _, blob in res], [r.nbytes for r in reads[a:b]], opts, sized=True)
    mine = [(a + i, code, zlib.crc32(blob), zlib.crc32(dec)) for i, ((code, blob), (_, dec)) in enumerate(zip(res, back))]
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)  # bookkeeping only: the data path has no collective
    dist.barrier()
    if rank == 0:
        out_q.put((gathered, [(zlib.crc32(r.tobytes()), r.nbytes) for r in reads]))
    dist.destroy_process_group()


def test_two_ranks_shard_by_read_index():
    import torch.multiprocessing as mp
    product_dir = os.path.join(ROOT, "tests")
    assert os.path.isdir(product_dir)
    from vbz_compression_b200 import build as vbuild
    vbuild.build_cpusim()  # once, before the ranks race for it
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    gathered, truth = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    seen = sorted(x for part in gathered for x in part)
    assert [i for i, *_ in seen] == list(range(len(truth)))  # every read exactly once
    for (i, code, _, crc_back), (crc_raw, nbytes) in zip(seen, truth):
        assert code < 0xFFFFFFF9 and crc_back == crc_raw, i
