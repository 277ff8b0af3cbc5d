"""What the host <-> device links of this box deliver when every GPU moves data at once (the ceiling of bench.py's `e2e` at N GPUs).
torchrun --nproc-per-node N tools/pcie_all.py: every rank copies 1 GiB pinned buffers host -> device, device -> host, and both at once,
all ranks at the same time; rank 0 prints per-direction GB/s per rank and summed over the ranks."""
import os
import time

import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1 << 30
h1 = torch.empty(n, dtype=torch.uint8).pin_memory(); h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d1 = torch.empty(n, dtype=torch.uint8, device="cuda"); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
flag = torch.zeros(1, device="cuda")


def barrier():
    if world > 1:
        dist.all_reduce(flag)
    torch.cuda.synchronize()


def timed(fn, reps=4):
    fn(); torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return reps * n / (time.perf_counter() - t0) / 1e9


def both():
    with torch.cuda.stream(s1):
        d1.copy_(h1, non_blocking=True)
    with torch.cuda.stream(s2):
        h2.copy_(d2, non_blocking=True)


res = [timed(lambda: d1.copy_(h1, non_blocking=True)), timed(lambda: h2.copy_(d2, non_blocking=True)), timed(both)]
t = torch.tensor(res, device="cuda", dtype=torch.float64)
if world > 1:
    all_t = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(all_t, t)
else:
    all_t = [t]
if rank == 0:
    rows = [x.cpu().tolist() for x in all_t]
    for name, k in (("host -> device", 0), ("device -> host", 1), ("both at once, each way", 2)):
        print("%-24s per rank %s  sum %.1f GB/s" % (name, " ".join("%.1f" % r[k] for r in rows), sum(r[k] for r in rows)))
    print("cpus %s, ranks %d" % (os.cpu_count(), world))
if world > 1:
    dist.destroy_process_group()
