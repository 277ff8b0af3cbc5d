#!/usr/bin/env python3
"""bench.py - the driver-facing benchmark of the VBZ hot path on B200.

    python bench.py --gpus N --steps K --warmup W                 (N > 1: launched under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W

Workload (default, BASELINE.json configs[2]): full VBZ {zigzag, int16, zstd 1, v1} - the options
the reference's users run (pyvbz and the H5Z filter default to level 1) and the path BASELINE's north_star
describes (stage 1 + zstd, compression ratio next to the reference's) - on synthetic nanopore-like int16
reads of 10k-1M samples (lengths floor(exp(U[ln 1e4, ln 1e6])), numpy default_rng(20261017 + rank)), generated
on the device: 40,000 reads (~17 GB raw) per GPU, the most that leaves room for the four batch-sized buffers
of the two-stage path.  One STEP = encode the whole batch, then decode it, inputs resident in HBM (`value`).
The same run also times the transform stage alone (configs[1], zstd level 0) on the same batch and reports it
as `transform_stage`, with its own HBM roofline; `--workload transform` makes that the whole benchmark
(100,000 reads, ~43 GB raw per GPU).

Metric: raw-sample GB/s = raw bytes that went through the codec (encode pass + decode pass) / time.
Weak scaling: every rank processes its own batch of the same shape; no data-path collective.

Printed JSON line: see the contract in the task description; extra keys `encode_gbs`, `decode_gbs`,
`ratio`, `roofline` (dominant kernel group + every stage), `transform_stage`, `cpu_baseline`, `e2e`, `clocks`, `gpu_launches`.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SEED = 20261017
REF_DIR = os.path.join(ROOT, "oracle", "_ref")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="full", choices=["transform", "full"])
    ap.add_argument("--reads", type=int, default=0, help="reads per GPU (default: 40,000 for the full workload, whose two-stage pipeline keeps "
                    "a fourth batch-sized buffer in HBM; 100,000 for --workload transform)")
    ap.add_argument("--cpu-sample-gib", type=float, default=1.0, help="raw GiB of the workload timed on the host cores")
    ap.add_argument("--e2e-sample-gib", type=float, default=2.0, help="raw GiB of the workload timed end to end with host buffers")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--verify", action="store_true", default=True)
    return ap.parse_args()


def options_for(workload):
    from vbz_compression_b200 import CompressionOptions
    return CompressionOptions(True, 2, 0 if workload == "transform" else 1, 0 if workload == "transform" else 1)


def read_lengths(n_reads, seed):
    rng = np.random.default_rng(seed)
    return np.floor(np.exp(rng.uniform(np.log(1e4), np.log(1e6), n_reads))).astype(np.int64)


# --------------------------------------------------------------------------------------------------
# synthetic signal
# --------------------------------------------------------------------------------------------------
def nanopore_like_numpy(n, rng):
    change = rng.random(n) < 0.1
    change[0] = True
    seg = np.cumsum(change) - 1
    levels = rng.normal(400.0, 90.0, int(seg[-1]) + 1)
    x = levels[seg] + rng.normal(0.0, 9.0, n)
    return np.clip(np.rint(x), -4096, 4095).astype(np.int16)


def nanopore_like_torch(out, seed):
    """Fill the int16 device tensor `out` (any length) chunk by chunk: level ~ N(400, 90) held for
    Geometric(0.1) samples, sample = clip(level + N(0, 9), -4096, 4095)."""
    import torch
    g = torch.Generator(device=out.device)
    g.manual_seed(seed)
    n = out.numel()
    chunk = 1 << 27
    for s in range(0, n, chunk):
        m = min(chunk, n - s)
        change = torch.rand(m, device=out.device, generator=g) < 0.1
        change[0] = True
        seg = torch.cumsum(change.to(torch.int32), 0) - 1
        nseg = int(seg[-1].item()) + 1
        levels = torch.randn(nseg, device=out.device, generator=g) * 90.0 + 400.0
        x = levels[seg.long()] + torch.randn(m, device=out.device, generator=g) * 9.0
        out[s:s + m] = torch.clamp(torch.round(x), -4096, 4095).to(torch.int16)
        del change, seg, levels, x


# --------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi while the timed region runs)
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = sorted(sm)[len(sm) // 2:]  # upper half = samples under load
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------------
# CPU baseline: the compiled reference on the host cores, bounded sample
# --------------------------------------------------------------------------------------------------
def run_cpu_bench(raw, lens_bytes, opts, reps=3, threads=0):
    exe = os.path.join(REF_DIR, "cpu_bench")
    lib = os.path.join(REF_DIR, "libvbz_ref.so")
    if not (os.path.exists(exe) and os.path.exists(lib)):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "all"], check=False)
    if not (os.path.exists(exe) and os.path.exists(lib)):
        return None
    with tempfile.TemporaryDirectory() as td:
        rp, lp = os.path.join(td, "raw.bin"), os.path.join(td, "lens.i64")
        raw.tofile(rp)
        np.asarray(lens_bytes, np.int64).tofile(lp)
        out = subprocess.run([exe, lib, rp, lp, str(int(opts.perform_delta_zig_zag)), str(opts.integer_size), str(opts.zstd_compression_level),
                              str(opts.vbz_version), str(threads), str(reps)], capture_output=True, text=True)
    if out.returncode != 0:
        sys.stderr.write("cpu_bench failed: " + out.stderr + out.stdout + "\n")
        return None
    return json.loads(out.stdout.strip().splitlines()[-1])


def cpu_sample(lens, gib, seed):
    """First reads of the workload's length sequence up to `gib` GiB raw, signal from numpy (same model)."""
    budget = int(gib * (1 << 30)) // 2
    take, tot = 0, 0
    while take < len(lens) and tot + lens[take] <= budget:
        tot += int(lens[take]); take += 1
    take = max(take, 1)
    rng = np.random.default_rng(seed + 7)
    sel = lens[:take]
    raw = nanopore_like_numpy(int(sel.sum()), rng)
    return raw, sel * 2


def combined_gbs(raw_bytes, enc_s, dec_s):
    return 2.0 * raw_bytes / (enc_s + dec_s) / 1e9


# --------------------------------------------------------------------------------------------------
def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    opts = options_for(args.workload)
    lens = read_lengths(args.reads, SEED)
    raw, lens_b = cpu_sample(lens, args.cpu_sample_gib, SEED)
    enc, dec, last = [], [], None
    total = args.warmup + args.steps
    for i in range(total):
        r = run_cpu_bench(raw, lens_b, opts, reps=1)
        if r is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/cpu_bench could not be built or run"}))
            return
        if i >= args.warmup:
            enc.append(r["encode_s"]); dec.append(r["decode_s"])
        last = r
    e, d = float(np.mean(enc)), float(np.mean(dec))
    val = combined_gbs(raw.nbytes, e, d)
    sample = "first %d reads (%.2f GiB raw) of the workload's length sequence, numpy signal of the same model; one vbz_*_sized call per read" % (len(lens_b), raw.nbytes / 2**30)
    line = {
        "impl": "reference", "metric": "raw int16 GB/s encode+decode", "value": val, "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": (e + d) * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int16", "data": "synthetic",
        "config": workload_config(args, opts, sample_note=sample),
        "encode_gbs": raw.nbytes / e / 1e9, "decode_gbs": raw.nbytes / d / 1e9, "ratio": last["ratio"],
        "cpu_baseline": {"value": val, "unit": "GB/s", "cores": last["threads"], "kind": "reference", "sample": sample},
        "e2e": {"value": val, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def default_reads(args):
    if args.reads <= 0:
        args.reads = 100_000 if args.workload == "transform" else 40_000
    return args


def workload_config(args, opts, sample_note=None):
    c = {"workload": "transform-only VBZ (delta+zigzag+StreamVByte, zstd level 0), int16" if args.workload == "transform"
         else "full VBZ int16 zigzag + zstd level 1 (v1)",
         "options": {"perform_delta_zig_zag": bool(opts.perform_delta_zig_zag), "integer_size": opts.integer_size,
                     "zstd_compression_level": opts.zstd_compression_level, "vbz_version": opts.vbz_version},
         "reads_per_gpu": args.reads, "read_length_samples": "floor(exp(U[ln 1e4, ln 1e6]))", "signal": "nanopore-like random walk (BASELINE.md section 3)",
         "step": "encode the whole batch, then decode it", "l2": "inputs (tens of GB) far larger than the 126 MB L2; no flush needed",
         "parallelism": "reads sharded by index, one process per GPU, no collective"}
    if sample_note:
        c["sample"] = sample_note
    return c


def main_ours(args):
    import torch
    import torch.distributed as dist
    from vbz_compression_b200 import Batch, load

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # NCCL logs (its version banner) go to stdout otherwise: stdout carries the JSON line only
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libvbz_b200 has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    lib = load()
    batch = Batch(local, lib=lib)
    opts = options_for(args.workload)

    # ---- the batch, resident in HBM ---------------------------------------------------------------------
    lens = read_lengths(args.reads, SEED + rank)
    n = len(lens)
    src_size = (lens * 2).astype(np.uint32)
    a16 = lambda v: (v + 15) & ~np.uint64(15)
    src_slot = a16(src_size.astype(np.uint64))
    src_off = np.concatenate([[0], np.cumsum(src_slot)[:-1]]).astype(np.uint64)
    caps = np.array([lib.vbz_b200_compress_bound(int(s), ctypes.byref(opts), 1) for s in src_size], np.uint64)
    dst_off = np.concatenate([[0], np.cumsum(caps)[:-1]]).astype(np.uint64)
    raw_bytes = int(src_size.astype(np.int64).sum())
    total_src, total_dst = int(src_slot.sum()), int(caps.sum())
    d_src = torch.empty(total_src, dtype=torch.uint8, device=dev)
    nanopore_like_torch(d_src.view(torch.int16), SEED + rank)
    d_cmp = torch.empty(total_dst, dtype=torch.uint8, device=dev)
    d_back = torch.empty(total_src, dtype=torch.uint8, device=dev)
    d_res_c = torch.zeros(n, dtype=torch.int32, device=dev)
    d_res_d = torch.zeros(n, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    caps32 = caps.astype(np.uint32)

    def encode():
        batch.compress_device(d_src.data_ptr(), src_off, src_size, d_cmp.data_ptr(), dst_off, caps32, d_res_c.data_ptr(), opts, True, stream)

    csize = [None]

    def decode():
        batch.decompress_device(d_cmp.data_ptr(), dst_off, csize[0], d_back.data_ptr(), src_off, src_size, d_res_d.data_ptr(), opts, True, stream)

    # first pass: sizes of the compressed reads (needed by decode), correctness of the round trip
    encode()
    torch.cuda.synchronize()
    res_c = d_res_c.cpu().numpy().view(np.uint32)
    assert (res_c < 0xFFFFFFF9).all(), "encode reported errors"
    csize[0] = res_c.copy()
    comp_bytes = int(res_c.astype(np.int64).sum())
    decode()
    torch.cuda.synchronize()
    res_d = d_res_d.cpu().numpy().view(np.uint32)
    assert (res_d == src_size).all(), "decode reported errors"
    if args.verify:
        # poison the output, decode again, and count differing bytes over the WHOLE batch: only the
        # never-written padding between reads (slots are 16-byte aligned) may differ
        d_back.copy_(d_src)
        d_back.add_(1)
        decode()
        torch.cuda.synchronize()
        mism = 0
        for s0 in range(0, total_src, 1 << 30):
            mism += int((d_src[s0:s0 + (1 << 30)] != d_back[s0:s0 + (1 << 30)]).sum().item())
        assert mism == total_src - raw_bytes, "round trip mismatch: %d differing bytes, %d padding bytes" % (mism, total_src - raw_bytes)

    for _ in range(args.warmup):
        encode(); decode()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        # nvidia-smi's own start-up takes driver locks that stall kernel launches for most of a second: begin timing only
        # once it is sampling
        t_wait = time.time()
        while len(sampler.lines) < 2 and time.time() - t_wait < 8.0:
            time.sleep(0.05)
    if world > 1:
        dist.barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * args.steps + 1)]
    l0 = batch.kernel_launches
    torch.cuda.synchronize()
    ev[0].record()
    for k in range(args.steps):
        encode(); ev[2 * k + 1].record()
        decode(); ev[2 * k + 2].record()
    torch.cuda.synchronize()
    launches = batch.kernel_launches - l0
    clocks = sampler.stop() if rank == 0 else None
    enc_ms = [ev[2 * k].elapsed_time(ev[2 * k + 1]) for k in range(args.steps)]
    dec_ms = [ev[2 * k + 1].elapsed_time(ev[2 * k + 2]) for k in range(args.steps)]
    total_ms = ev[0].elapsed_time(ev[-1])
    if os.environ.get("VBZ_BENCH_VERBOSE"):
        sys.stderr.write("rank %d encode ms %s decode ms %s\n" % (rank, [round(x, 2) for x in enc_ms], [round(x, 2) for x in dec_ms]))
    t = torch.tensor([total_ms, float(np.mean(enc_ms)), float(np.mean(dec_ms))], dtype=torch.float64, device=dev)
    b = torch.tensor([raw_bytes, comp_bytes], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(b, op=dist.ReduceOp.SUM)
    total_ms, enc_m, dec_m = [float(x) for x in t.tolist()]
    all_raw, all_comp = [float(x) for x in b.tolist()]
    value = 2.0 * all_raw * args.steps / (total_ms * 1e-3) / 1e9

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- per-stage times (CUDA events inside the library, on the launching stream) and the transform stage alone --------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "B200_PROFILING.md fallback 6650 GB/s (of fallback)"
    f1, f2 = ctypes.c_float(), ctypes.c_float()
    lib.vbz_b200_set_stage_timing(batch.ctx, 1)
    st = {"enc1": [], "enc2": [], "dec2": [], "dec1": []}
    for _ in range(3):
        encode()
        assert lib.vbz_b200_stage_times(batch.ctx, ctypes.byref(f1), ctypes.byref(f2)) == 0
        st["enc1"].append(f1.value); st["enc2"].append(f2.value)
        decode()
        assert lib.vbz_b200_stage_times(batch.ctx, ctypes.byref(f1), ctypes.byref(f2)) == 0
        st["dec1"].append(f1.value); st["dec2"].append(f2.value)
    lib.vbz_b200_set_stage_timing(batch.ctx, 0)
    st = {k: float(np.mean(v)) for k, v in st.items()}

    transform_stage = None
    s1_bytes = comp_bytes - 4 * n  # stage-1 bytes of the batch (control + data bytes)
    if args.workload == "full":
        # BASELINE configs[1] on the same resident batch: {zigzag, int16, zstd 0, v0}
        t_opts = options_for("transform")
        t_caps = np.array([lib.vbz_b200_compress_bound(int(s), ctypes.byref(t_opts), 1) for s in src_size], np.uint32)
        t_size = [None]

        def t_encode():
            batch.compress_device(d_src.data_ptr(), src_off, src_size, d_cmp.data_ptr(), dst_off, t_caps, d_res_c.data_ptr(), t_opts, True, stream)

        def t_decode():
            batch.decompress_device(d_cmp.data_ptr(), dst_off, t_size[0], d_back.data_ptr(), src_off, src_size, d_res_d.data_ptr(), t_opts, True, stream)

        t_encode()
        torch.cuda.synchronize()
        t_size[0] = d_res_c.cpu().numpy().view(np.uint32).copy()
        assert (t_size[0] < 0xFFFFFFF9).all()
        s1_bytes = int(t_size[0].astype(np.int64).sum()) - 4 * n
        d_back.zero_()
        t_decode()
        torch.cuda.synchronize()
        assert (d_res_d.cpu().numpy().view(np.uint32) == src_size).all()
        for i in (0, n // 3, n - 1):  # (the 16-byte padding between reads is never written: compare reads, not the buffer)
            o, z = int(src_off[i]), int(src_size[i])
            assert torch.equal(d_back[o:o + z], d_src[o:o + z]), "transform stage round trip mismatch"
        for _ in range(2):
            t_encode(); t_decode()
        tk = max(3, min(args.steps, 5))
        tev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * tk + 1)]
        torch.cuda.synchronize()
        tev[0].record()
        for k in range(tk):
            t_encode(); tev[2 * k + 1].record()
            t_decode(); tev[2 * k + 2].record()
        torch.cuda.synchronize()
        te = float(np.mean([tev[2 * k].elapsed_time(tev[2 * k + 1]) for k in range(tk)]))
        td = float(np.mean([tev[2 * k + 1].elapsed_time(tev[2 * k + 2]) for k in range(tk)]))
        t_alg = raw_bytes + s1_bytes
        transform_stage = {
            "workload": "transform-only VBZ (delta+zigzag+StreamVByte, zstd level 0), int16 - BASELINE configs[1] on the same resident batch (this rank)",
            "steps": tk, "value": combined_gbs(raw_bytes, te * 1e-3, td * 1e-3), "unit": "GB/s",
            "encode_gbs": raw_bytes / te / 1e6, "decode_gbs": raw_bytes / td / 1e6, "ratio": (s1_bytes + 4 * n) / raw_bytes,
            "roofline": {"bound": "hbm", "kernel": "svb16_decode_kernel" if td >= te else "svb16_encode_kernel", "peak": peak, "unit": "GB/s",
                         "algorithmic_bytes_per_launch": t_alg, "achieved": t_alg / max(te, td) / 1e6, "frac": t_alg / max(te, td) / 1e6 / peak,
                         "encode": {"achieved": t_alg / te / 1e6, "frac": t_alg / te / 1e6 / peak, "ms": te},
                         "decode": {"achieved": t_alg / td / 1e6, "frac": t_alg / td / 1e6 / peak, "ms": td}}}
        # back to the workload's own compressed batch for whatever follows
        encode()
        torch.cuda.synchronize()

    # ---- roofline: algorithmic bytes of each kernel group / its live time; the headline entry is the slowest group ----------
    # stage 1 moves the samples once and the control + data bytes once; the zstd stage moves the stage-1 bytes once and
    # the frames once.  Times: CUDA events on the launching stream around the group's launches
    # (descriptor upload + tables + look-back reset + the kernels).
    alg1 = raw_bytes + s1_bytes
    alg2 = s1_bytes + comp_bytes
    groups = {}
    if args.workload == "full":
        groups = {"svb16_encode_kernel": (alg1, st["enc1"]), "zstd_encode_kernel": (alg2, st["enc2"]),
                  "zstd_units_kernel (decode)": (alg2, st["dec2"]), "svb16_decode_kernel": (alg1, st["dec1"])}
    else:
        groups = {"svb16_encode_kernel": (alg1, float(np.mean(enc_ms))), "svb16_decode_kernel": (alg1, float(np.mean(dec_ms)))}
    dom = max(groups, key=lambda k: groups[k][1])
    stages = {k: {"algorithmic_bytes_per_launch": b_, "ms": ms_, "achieved": b_ / ms_ / 1e6, "frac": b_ / ms_ / 1e6 / peak} for k, (b_, ms_) in groups.items()}
    traffic = None
    try:  # DRAM bytes per algorithmic byte of the dominant kernel from the committed `ncu --set full` capture, scaled to this launch
        tj = json.load(open(os.path.join(ROOT, "profiles", "dram_traffic.json")))
        key = dom.split(" ")[0]
        if key in tj:
            traffic = tj[key]["dram_bytes_per_algorithmic_byte"] * groups[dom][0]
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dom, "achieved": stages[dom]["achieved"], "peak": peak, "unit": "GB/s", "frac": stages[dom]["frac"],
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": groups[dom][0], "stages": stages,
                "note": "dominant = the kernel group with the largest share of the step; time = CUDA events on the launching stream around the "
                        "group's launches (descriptor upload + tables + look-back reset + the kernels); the entropy stage is issue-bound integer "
                        "work, its HBM fraction is reported for completeness; the HBM-bound kernels are the two svb16 ones (see transform_stage)"}

    # ---- CPU baseline (bounded sample of the same workload on the host cores) ---------------------------------
    cpu = None
    if not args.no_cpu_baseline:
        take, tot = 0, 0
        budget = int(args.cpu_sample_gib * (1 << 30))
        while take < n and tot + int(src_size[take]) <= budget:
            tot += int(src_size[take]); take += 1
        take = max(take, 1)
        host = np.concatenate([d_src[int(src_off[i]):int(src_off[i]) + int(src_size[i])].cpu().numpy() for i in range(take)])
        r = run_cpu_bench(host, src_size[:take].astype(np.int64), opts, reps=3)
        if r:
            cpu = {"value": combined_gbs(r["raw_bytes"], r["encode_s"], r["decode_s"]), "unit": "GB/s", "cores": r["threads"], "kind": "reference",
                   "encode_gbs": r["encode_gbs"], "decode_gbs": r["decode_gbs"], "ratio": r["ratio"],
                   "sample": "first %d reads (%.2f GiB raw) of this run's batch, copied to the host; oracle/_ref (unmodified reference, -O3 -mssse3, libzstd 1.5.5), one vbz_*_sized call per read, best of 3" % (take, r["raw_bytes"] / 2**30)}

    # ---- end to end through the host-pointer C-ABI (H2D + D2H inside the timed region) ---------------------------
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, batch, opts, d_src, src_off, src_size, n)

    line = {
        "metric": "raw int16 GB/s encode+decode", "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int16", "data": "synthetic",
        "config": workload_config(args, opts),
        "encode_gbs": all_raw / (enc_m * 1e-3) / 1e9, "decode_gbs": all_raw / (dec_m * 1e-3) / 1e9, "ratio": all_comp / all_raw,
        "raw_bytes_per_gpu": raw_bytes, "roofline": roofline, "transform_stage": transform_stage, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks, "gpu_launches": int(launches),
    }
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_e2e(args, batch, opts, d_src, src_off, src_size, n):
    """Same metric through vbz_b200_compress_batch / vbz_b200_decompress_batch with HOST buffers."""
    import torch
    take, tot = 0, 0
    budget = int(args.e2e_sample_gib * (1 << 30))
    while take < n and tot + int(src_size[take]) <= budget:
        tot += int(src_size[take]); take += 1
    take = max(take, 1)
    lib = batch.lib
    host_in = torch.empty(tot, dtype=torch.uint8).pin_memory()
    offs, p = [], 0
    for i in range(take):
        s = int(src_size[i])
        host_in[p:p + s].copy_(d_src[int(src_off[i]):int(src_off[i]) + s])
        offs.append(p); p += s
    caps = [int(lib.vbz_max_compressed_size(int(src_size[i]), ctypes.byref(opts))) for i in range(take)]
    coffs = np.concatenate([[0], np.cumsum(caps)[:-1]]).astype(np.int64)
    host_c = torch.empty(int(sum(caps)), dtype=torch.uint8).pin_memory()
    host_back = torch.empty(tot, dtype=torch.uint8).pin_memory()
    vp, u32 = ctypes.c_void_p, ctypes.c_uint32
    srcp = (vp * take)(*[host_in.data_ptr() + o for o in offs])
    cp = (vp * take)(*[host_c.data_ptr() + int(o) for o in coffs])
    bp = (vp * take)(*[host_back.data_ptr() + o for o in offs])
    ssz = (u32 * take)(*[int(src_size[i]) for i in range(take)])
    ccap = (u32 * take)(*caps)
    res_c = (u32 * take)()
    res_d = (u32 * take)()

    def step():
        rc = lib.vbz_b200_compress_batch(batch.ctx, srcp, ssz, cp, ccap, res_c, take, ctypes.byref(opts), 1)
        assert rc == 0, lib.vbz_b200_last_error()
        rc = lib.vbz_b200_decompress_batch(batch.ctx, cp, res_c, bp, ssz, res_d, take, ctypes.byref(opts), 1)
        assert rc == 0, lib.vbz_b200_last_error()

    step()
    assert torch.equal(host_in, host_back), "e2e round trip mismatch"
    comp = int(np.sum(np.frombuffer(res_c, np.uint32).astype(np.int64)))
    reps = 3
    t0 = time.perf_counter()
    for _ in range(reps):
        step()
    dt = (time.perf_counter() - t0) / reps
    return {"value": 2.0 * tot / dt / 1e9, "unit": "GB/s", "h2d_bytes_per_step": tot + comp, "d2h_bytes_per_step": comp + tot,
            "ms_per_step": dt * 1e3,
            "sample": "first %d reads (%.2f GiB raw) of the batch as pinned HOST buffers through vbz_b200_compress_batch + vbz_b200_decompress_batch (wall clock, %d reps)" % (take, tot / 2**30, reps)}


if __name__ == "__main__":
    a = default_reads(parse_args())
    if a.impl == "reference":
        main_reference(a)
    else:
        main_ours(a)
