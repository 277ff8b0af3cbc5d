#!/usr/bin/env python3
"""bench.py - the driver-facing benchmark of the VBZ hot path on B200.

    python bench.py --gpus N --steps K --warmup W                 (N > 1: launched under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W
    python bench.py --workload {full,transform,uint32,int8,tiled,h5z-decode} ...

Workloads (BASELINE.json configs):
  full        (default, configs[2]) full VBZ {zigzag, int16, zstd 1, v1} - the options the reference's users run (pyvbz and
              the H5Z filter default to level 1) and the path north_star describes - on synthetic nanopore-like int16 reads
              of 10k-1M samples (lengths floor(exp(U[ln 1e4, ln 1e6])), numpy default_rng(20261017 + rank)), generated on
              the device: 100,000 reads (~43 GB raw) per GPU.  One STEP = encode the whole batch, then decode it, inputs
              resident in HBM (`value`).  The same run also times the transform stage alone (configs[1], zstd level 0) on
              the same batch (`transform_stage`).
  transform   (configs[1]) the transform stage as the whole benchmark.
  uint32      (configs[3]) {no zigzag, 4-byte integers, zstd 3, v0}: values U[0, 2^20), reads of half as many values.
  int8        {zigzag, 1-byte integers, zstd 1, v1}: the nibble-packed path of vbz/v1; the signal >> 3, reads of twice as many values.
  tiled       the full options on the reference benchmark's kind of input (vbz/perf/test_data_generator.h): reads of U[30000, 200000]
              samples that repeat one 15,643-sample trace - where the zstd stage's matches are everything.
  h5z-decode  (configs[4]) decode-only: 1,000,000 reads presented as H5Z chunk buffers `[u32 size][zstd frame]` WRITTEN BY THE
              REFERENCE (oracle/_ref, outside the timed region), sharded by read index over the ranks with byte-balanced
              contiguous ranges (vbz_compression_b200/shard.py), decoded through the batch form of the filter 32020 callback
              with cd_values (0,2,1,1) and (1,2,1,1).  STRONG scaling: the job is fixed, every rank decodes its range.

Metric: raw-sample GB/s = raw bytes that went through the codec (encode pass + decode pass; decode pass only for
h5z-decode) / time.  `value`: batch resident in HBM, CUDA events.  `e2e`: the same through the host-pointer C-ABI with
pinned HOST buffers, H2D + D2H inside the timed region, measured on EVERY rank at the same time and aggregated
(bytes summed, time = max over ranks).

Printed JSON line: see the contract in the task description; extra keys `encode_gbs`, `decode_gbs`, `ratio`, `roofline`
(dominant kernel group + every stage), `transform_stage`, `cpu_baseline`, `e2e`, `clocks`, `gpu_launches`.
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
WORKLOADS = ("full", "transform", "uint32", "int8", "tiled", "h5z-decode")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="full", choices=WORKLOADS)
    ap.add_argument("--reads", type=int, default=0, help="reads per GPU (default 100,000); h5z-decode: reads of the whole job (default 1,000,000)")
    ap.add_argument("--cpu-sample-gib", type=float, default=1.0, help="raw GiB of the workload timed on the host cores")
    ap.add_argument("--e2e-sample-gib", type=float, default=2.0, help="raw GiB of the workload timed end to end with host buffers (per rank)")
    ap.add_argument("--pool-gib", type=float, default=2.0, help="h5z-decode: raw GiB of distinct reference-written chunks (read i of the job = chunk i mod pool)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-verify", action="store_true")
    return ap.parse_args()


# --------------------------------------------------------------------------------------------------
# workloads
# --------------------------------------------------------------------------------------------------
def options_for(workload):
    from vbz_compression_b200 import CompressionOptions
    if workload == "transform":
        return CompressionOptions(True, 2, 0, 0)
    if workload == "uint32":
        return CompressionOptions(False, 4, 3, 0)
    if workload == "int8":
        return CompressionOptions(True, 1, 1, 1)
    return CompressionOptions(True, 2, 1, 1)


def workload_name(workload):
    return {"full": "full VBZ int16 zigzag + zstd level 1 (v1)",
            "transform": "transform-only VBZ (delta+zigzag+StreamVByte, zstd level 0), int16",
            "uint32": "uint32 no-zigzag + zstd level 3 (4-byte integer path, v0)",
            "int8": "int8 zigzag + zstd level 1, v1: the nibble-packed 1-byte path (ref vbz/v1/vbz_streamvbyte_impl.h:20-301)",
            "tiled": "full VBZ int16 zigzag + zstd level 1 (v1) on the reference benchmark's kind of input: reads that tile one 15,643-sample trace (ref vbz/perf/test_data_generator.h:29-74)",
            "h5z-decode": "decode-only: reference-written H5Z chunk buffers (int16 zigzag zstd 1, cd_values (0,2,1,1) and (1,2,1,1)) sharded by read index"}[workload]


def read_lengths(n_reads, seed, workload="full"):
    rng = np.random.default_rng(seed)
    if workload == "tiled":  # the reference's SignalGenerator: U[30000, 200000] samples per read
        return rng.integers(30000, 200001, n_reads).astype(np.int64)
    return np.floor(np.exp(rng.uniform(np.log(1e4), np.log(1e6), n_reads))).astype(np.int64)


TILE_SAMPLES = 15643  # length of the trace the reference's benchmark tiles (vbz/test/test_data.h)


def tile_trace():
    """the trace the `tiled` workload repeats: nanopore-like, as long as the reference's (its samples stay where they are: test vectors, not bench input)"""
    return nanopore_like_numpy(TILE_SAMPLES, np.random.default_rng(5))


def value_counts(workload, lens):
    """values per read: the uint32 workload uses half as many (4-byte) values, so a read has the same bytes"""
    return lens // 2 if workload == "uint32" else (lens * 2 if workload == "int8" else lens)


def default_reads(args):
    if args.reads <= 0:
        # (int8: the output slots of the device-resident API are sized for the worst case of the 1-byte path, 4.25 bytes per value: 25,000 reads fit)
        args.reads = 1_000_000 if args.workload == "h5z-decode" else (25_000 if args.workload == "int8" else 100_000)
    return args


def workload_config(args, opts):
    c = {"workload": workload_name(args.workload),
         "options": {"perform_delta_zig_zag": bool(opts.perform_delta_zig_zag), "integer_size": opts.integer_size,
                     "zstd_compression_level": opts.zstd_compression_level, "vbz_version": opts.vbz_version},
         "read_length_samples": "U[30000, 200000]" if args.workload == "tiled" else
                                "floor(exp(U[ln 1e4, ln 1e6]))" + (" / 2 values of 4 bytes" if args.workload == "uint32" else " x 2 values of 1 byte" if args.workload == "int8" else ""),
         "signal": {"uint32": "uniform U[0, 2^20) uint32", "int8": "nanopore-like random walk (BASELINE.md section 3) >> 3, as int8",
                    "tiled": "every read = one nanopore-like trace of 15,643 samples repeated from its first sample on"}.get(args.workload, "nanopore-like random walk (BASELINE.md section 3)"),
         "l2": "inputs (GBs per launch) far larger than the 126 MB L2; no flush needed",
         "cpu_arm_sample_gib": args.cpu_sample_gib}
    if args.workload == "h5z-decode":
        c.update({"reads_total": args.reads, "pool_gib": args.pool_gib, "step": "decode every read of the job once",
                  "parallelism": "reads sharded by index over the ranks (byte-balanced contiguous ranges), one process per GPU, no collective"})
    else:
        c.update({"reads_per_gpu": args.reads, "step": "encode the whole batch, then decode it",
                  "parallelism": "reads sharded by index, one process per GPU, no collective"})
    return c


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


def signal_numpy(workload, lens, rng):
    """host-side signal of the workload's model for reads of `lens` values, back to back, as bytes"""
    n_values = int(np.sum(lens))
    if workload == "uint32":
        return rng.integers(0, 1 << 20, n_values, dtype=np.uint32).view(np.uint8)
    if workload == "int8":
        return (nanopore_like_numpy(n_values, rng) >> 3).astype(np.int8).view(np.uint8)
    if workload == "tiled":
        t = tile_trace()
        return np.concatenate([np.resize(t, int(n)) for n in lens]).view(np.uint8)
    return nanopore_like_numpy(n_values, rng).view(np.uint8)


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


def fill_device_signal(workload, d_src, seed, src_off=None, lens=None):
    import torch
    if workload == "tiled":  # every read starts at the trace's first sample
        longest = int(max(lens))
        t = torch.from_numpy(np.resize(tile_trace(), longest)).to(d_src.device)
        v = d_src.view(torch.int16)
        v.zero_()
        for o, n in zip(src_off, lens):
            v[int(o) // 2:int(o) // 2 + int(n)] = t[:int(n)]
        return
    if workload == "int8":  # the int16 model >> 3, piece by piece
        v8 = d_src.view(torch.int8)
        piece = 1 << 28
        tmp = torch.empty(min(piece, v8.numel()), dtype=torch.int16, device=d_src.device)
        for k, s0 in enumerate(range(0, v8.numel(), piece)):
            m = min(piece, v8.numel() - s0)
            nanopore_like_torch(tmp[:m], seed + 1000 * k)
            v8[s0:s0 + m] = (tmp[:m] >> 3).to(torch.int8)
        del tmp
        return
    if workload == "uint32":
        g = torch.Generator(device=d_src.device)
        g.manual_seed(seed)
        v = d_src.view(torch.int32)
        chunk = 1 << 28
        for s in range(0, v.numel(), chunk):
            m = min(chunk, v.numel() - s)
            v[s:s + m] = torch.randint(0, 1 << 20, (m,), device=d_src.device, generator=g, dtype=torch.int32)
    else:
        nanopore_like_torch(d_src.view(torch.int16), seed)


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

    def start_and_wait(self):
        """nvidia-smi's own start-up takes driver locks that stall kernel launches for most of a second: begin timing only
        once it is sampling"""
        self.start()
        t_wait = time.time()
        while self.proc and len(self.lines) < 2 and time.time() - t_wait < 8.0:
            time.sleep(0.05)

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


def pin_to_gpu_numa(local):
    """Run this rank (and the pinned staging it allocates: first touch) on the cores next to its GPU."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        cpus = open("/sys/bus/pci/devices/%s/local_cpulist" % bdf).read().strip()
        node = open("/sys/bus/pci/devices/%s/numa_node" % bdf).read().strip()
        ids = set()
        for part in cpus.split(","):
            if "-" in part:
                a, b = part.split("-"); ids.update(range(int(a), int(b) + 1))
            elif part:
                ids.add(int(part))
        ids &= set(os.sched_getaffinity(0))
        if ids:
            os.sched_setaffinity(0, ids)
        return {"pci": bdf, "numa_node": node, "cpus": cpus}
    except Exception as e:  # noqa: BLE001
        return {"unavailable": str(e)[:80]}


def measured_peak():
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "B200_PROFILING.md fallback 6650 GB/s (of fallback)"
    return peak, src


# --------------------------------------------------------------------------------------------------
# CPU arm: the compiled reference on the host cores, bounded sample
# --------------------------------------------------------------------------------------------------
def _ensure_ref():
    exe = os.path.join(REF_DIR, "cpu_bench")
    lib = os.path.join(REF_DIR, "libvbz_ref.so")
    if not (os.path.exists(exe) and os.path.exists(lib)):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "all"], check=False)
    return (exe, lib) if os.path.exists(exe) and os.path.exists(lib) else (None, None)


def _opt_args(opts):
    return [str(int(opts.perform_delta_zig_zag)), str(opts.integer_size), str(opts.zstd_compression_level), str(opts.vbz_version)]


def run_cpu_bench(raw, lens_bytes, opts, reps=3, threads=0, dump_prefix=None):
    """encode + decode of `raw` (reads back to back) with the reference on the host cores"""
    exe, lib = _ensure_ref()
    if not exe:
        return None
    with tempfile.TemporaryDirectory() as td:
        rp, lp = os.path.join(td, "raw.bin"), os.path.join(td, "lens.i64")
        np.asarray(raw).tofile(rp)
        np.asarray(lens_bytes, np.int64).tofile(lp)
        cmd = [exe, lib, rp, lp] + _opt_args(opts) + [str(threads), str(reps)] + ([dump_prefix] if dump_prefix else [])
        out = subprocess.run(cmd, capture_output=True, text=True)
    if out.returncode != 0:
        sys.stderr.write("cpu_bench failed: " + out.stderr + out.stdout + "\n")
        return None
    return json.loads(out.stdout.strip().splitlines()[-1])


def run_cpu_decode(chunks, csizes, raw, lens_bytes, opts, reps=3, threads=0):
    """decode-only: `chunks` ([u32][payload] back to back, whoever wrote them) with the reference; compares with raw"""
    exe, lib = _ensure_ref()
    if not exe:
        return None
    with tempfile.TemporaryDirectory() as td:
        cp, sp, rp, lp = [os.path.join(td, f) for f in ("chunks.bin", "csizes.u32", "raw.bin", "lens.i64")]
        np.asarray(chunks).tofile(cp)
        np.asarray(csizes, np.uint32).tofile(sp)
        np.asarray(raw).tofile(rp)
        np.asarray(lens_bytes, np.int64).tofile(lp)
        out = subprocess.run([exe, lib, "--decode", cp, sp, rp, lp] + _opt_args(opts) + [str(threads), str(reps)], capture_output=True, text=True)
    if out.returncode not in (0, 1):
        sys.stderr.write("cpu_bench --decode failed: " + out.stderr + out.stdout + "\n")
        return None
    return json.loads(out.stdout.strip().splitlines()[-1])


def frame_content_size(h):
    """content size field of a zstd frame header (RFC 8878 3.1.1.1); h = the first bytes of the frame"""
    assert bytes(h[:4]) == b"\x28\xb5\x2f\xfd"
    fhd = int(h[4])
    single, flag = (fhd >> 5) & 1, fhd >> 6
    p = 5 + (0 if single else 1) + (0, 1, 2, 4)[fhd & 3]
    nb = (1 if single else 0, 2, 4, 8)[flag]
    v = int.from_bytes(bytes(h[p:p + nb]), "little")
    return v + 256 if nb == 2 else v


def sample_prefix(sizes, budget_bytes):
    take, tot = 0, 0
    while take < len(sizes) and tot + int(sizes[take]) <= budget_bytes:
        tot += int(sizes[take]); take += 1
    if take == 0:
        take, tot = 1, int(sizes[0])
    return take, tot


def combined_gbs(raw_bytes, enc_s, dec_s):
    return 2.0 * raw_bytes / (enc_s + dec_s) / 1e9


def reference_pool(args, opts, seed):
    """-> (raw uint8 array, byte count per read, chunk bytes, chunk sizes): the first reads of the length sequence up to
    --pool-gib raw, signal from numpy, compressed by the unmodified reference (one vbz_compress_sized per read)"""
    lens = read_lengths(200_000, seed)
    take, _ = sample_prefix(lens * 2, int(args.pool_gib * (1 << 30)))
    rng = np.random.default_rng(seed + 7)
    sel = lens[:take]
    raw = nanopore_like_numpy(int(sel.sum()), rng).view(np.uint8)
    with tempfile.TemporaryDirectory() as td:
        pre = os.path.join(td, "pool")
        r = run_cpu_bench(raw, sel * 2, opts, reps=1, dump_prefix=pre)
        if r is None:
            return None
        chunks = np.fromfile(pre + ".bin", np.uint8)
        csz = np.fromfile(pre + ".sizes.u32", np.uint32)
    return raw, (sel * 2).astype(np.int64), chunks, csz


# --------------------------------------------------------------------------------------------------
def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    opts = options_for(args.workload)
    total = args.warmup + args.steps
    if args.workload == "h5z-decode":
        pool = reference_pool(args, opts, SEED)
        if pool is None:
            emit_line(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/cpu_bench could not be built or run"}))
            return
        raw, lens_b, chunks, csz = pool
        take, tot = sample_prefix(lens_b, int(args.cpu_sample_gib * (1 << 30)))
        ctake = int(csz[:take].astype(np.int64).sum())
        dec, last = [], None
        for i in range(total):
            r = run_cpu_decode(chunks[:ctake], csz[:take], raw[:tot], lens_b[:take], opts, reps=1)
            if r is None or r["errors"]:
                emit_line(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/cpu_bench --decode failed"}))
                return
            if i >= args.warmup:
                dec.append(r["decode_s"])
            last = r
        d = float(np.mean(dec))
        val = tot / d / 1e9
        sample = "first %d chunks (%.2f GiB raw) of the pool, decoded by the reference, one vbz_decompress_sized call per chunk, all host cores" % (take, tot / 2**30)
        line = {"impl": "reference", "metric": "raw int16 GB/s decode", "value": val, "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": d * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int16",
                "data": "synthetic", "config": workload_config(args, opts), "decode_gbs": val, "ratio": last["ratio"],
                "cpu_baseline": {"value": val, "unit": "GB/s", "cores": last["threads"], "kind": "reference", "sample": sample},
                "e2e": {"value": val, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        emit_line(json.dumps(line))
        return
    lens = value_counts(args.workload, read_lengths(args.reads, SEED, args.workload))
    isz = opts.integer_size
    take, tot = sample_prefix(lens * isz, int(args.cpu_sample_gib * (1 << 30)))
    rng = np.random.default_rng(SEED + 7)
    raw = signal_numpy(args.workload, lens[:take], rng)
    lens_b = lens[:take] * isz
    enc, dec, last = [], [], None
    for i in range(total):
        r = run_cpu_bench(raw, lens_b, opts, reps=1)
        if r is None:
            emit_line(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/cpu_bench could not be built or run"}))
            return
        if i >= args.warmup:
            enc.append(r["encode_s"]); dec.append(r["decode_s"])
        last = r
    e, d = float(np.mean(enc)), float(np.mean(dec))
    val = combined_gbs(raw.nbytes, e, d)
    sample = "first %d reads (%.2f GiB raw) of the workload's length sequence, numpy signal of the same model; one vbz_*_sized call per read, all host cores" % (len(lens_b), raw.nbytes / 2**30)
    line = {
        "impl": "reference", "metric": "raw int16 GB/s encode+decode", "value": val, "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": (e + d) * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"uint32": "uint32", "int8": "int8"}.get(args.workload, "int16"), "data": "synthetic",
        "config": workload_config(args, opts),
        "encode_gbs": raw.nbytes / e / 1e9, "decode_gbs": raw.nbytes / d / 1e9, "ratio": last["ratio"],
        "cpu_baseline": {"value": val, "unit": "GB/s", "cores": last["threads"], "kind": "reference", "sample": sample},
        "e2e": {"value": val, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit_line(json.dumps(line))


# --------------------------------------------------------------------------------------------------
class Dist:
    """torch.distributed plumbing: barrier + the two reductions the timing needs.  No data-path collective."""

    def __init__(self):
        import torch
        self.torch = torch
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: libvbz_b200 has no CPU path")
        if self.world > 1:
            import torch.distributed as dist
            self.dist = dist
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # NCCL logs (its version banner) go to stdout otherwise: stdout carries the JSON line only
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()

    def reduce(self, values, op):
        t = self.torch.tensor(values, dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX if op == "max" else self.dist.ReduceOp.SUM)
        return [float(x) for x in t.tolist()]

    def close(self):
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()


def stage_roofline(groups, peak, peak_src, note):
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
    return {"bound": "hbm", "kernel": dom, "achieved": stages[dom]["achieved"], "peak": peak, "unit": "GB/s", "frac": stages[dom]["frac"],
            "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": groups[dom][0], "stages": stages, "note": note}


def main_codec(args):
    """full / transform / uint32: encode the resident batch, then decode it"""
    import torch
    from vbz_compression_b200 import Batch, load

    D = Dist()
    world, rank, local, dev = D.world, D.rank, D.local, D.dev
    numa = pin_to_gpu_numa(local)
    lib = load()
    batch = Batch(local, lib=lib)
    opts = options_for(args.workload)
    isz = opts.integer_size
    two_stage = opts.zstd_compression_level != 0
    s1_name = {"uint32": "svb<4,nozz>", "int8": "svb_half<zz>"}.get(args.workload, "svb16")

    # ---- the batch, resident in HBM ---------------------------------------------------------------------
    lens = value_counts(args.workload, read_lengths(args.reads, SEED + rank, args.workload))
    n = len(lens)
    src_size = (lens * isz).astype(np.uint32)
    a16 = lambda v: (v + 15) & ~np.uint64(15)
    src_slot = a16(src_size.astype(np.uint64))
    src_off = np.concatenate([[0], np.cumsum(src_slot)[:-1]]).astype(np.uint64)
    caps = np.array([lib.vbz_b200_compress_bound(int(s), ctypes.byref(opts), 1) for s in src_size], np.uint64)
    dst_off = np.concatenate([[0], np.cumsum(caps)[:-1]]).astype(np.uint64)
    raw_bytes = int(src_size.astype(np.int64).sum())
    total_src, total_dst = int(src_slot.sum()), int(caps.sum())
    d_src = torch.empty(total_src, dtype=torch.uint8, device=dev)
    fill_device_signal(args.workload, d_src, SEED + rank, src_off, lens)
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
    if not args.no_verify:
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
    D.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start_and_wait()
    D.barrier()
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
    total_ms, enc_m, dec_m = D.reduce([total_ms, float(np.mean(enc_ms)), float(np.mean(dec_ms))], "max")
    all_raw, all_comp, all_launches = D.reduce([raw_bytes, comp_bytes, launches], "sum")
    value = 2.0 * all_raw * args.steps / (total_ms * 1e-3) / 1e9

    # ---- end to end through the host-pointer C-ABI (H2D + D2H inside the timed region), every rank at the same time ---------
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, D, batch, opts, d_src, src_off, src_size, n)
    if rank != 0:
        D.close()
        return

    # ---- per-stage times (CUDA events inside the library, on the launching stream) and the transform stage alone --------
    peak, peak_src = measured_peak()
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
    if two_stage:
        # BASELINE configs[1] on the same resident batch: the same options with zstd level 0
        from vbz_compression_b200 import CompressionOptions
        t_opts = CompressionOptions(bool(opts.perform_delta_zig_zag), isz, 0, opts.vbz_version)
        t_caps = np.array([lib.vbz_b200_compress_bound(int(s), ctypes.byref(t_opts), 1) for s in src_size], np.uint32)
        assert (t_caps.astype(np.uint64) <= caps).all()
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
            "workload": "transform-only (zstd level 0) on the same resident batch (this rank) - BASELINE configs[1] for int16",
            "steps": tk, "value": combined_gbs(raw_bytes, te * 1e-3, td * 1e-3), "unit": "GB/s",
            "encode_gbs": raw_bytes / te / 1e6, "decode_gbs": raw_bytes / td / 1e6, "ratio": (s1_bytes + 4 * n) / raw_bytes,
            "roofline": {"bound": "hbm", "kernel": s1_name + ("_decode_kernel" if td >= te else "_encode_kernel"), "peak": peak, "unit": "GB/s",
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
    if two_stage:
        groups = {s1_name + "_encode_kernel": (alg1, st["enc1"]), "zstd_encode_kernel": (alg2, st["enc2"]),
                  "zstd_units_kernel (decode)": (alg2, st["dec2"]), s1_name + "_decode_kernel": (alg1, st["dec1"])}
    else:
        groups = {s1_name + "_encode_kernel": (alg1, float(np.mean(enc_ms))), s1_name + "_decode_kernel": (alg1, float(np.mean(dec_ms)))}
    roofline = stage_roofline(groups, peak, peak_src,
                              "dominant = the kernel group with the largest share of the step; time = CUDA events on the launching stream around the "
                              "group's launches (descriptor upload + tables + look-back reset + the kernels); the entropy stage is issue-bound integer "
                              "work, its HBM fraction is reported for completeness; the HBM-bound kernels are the two stage-1 ones (see transform_stage)")

    # ---- CPU baseline (bounded sample of the same workload on the host cores) + cross-decode of GPU-written frames ---------
    cpu = None
    if not args.no_cpu_baseline:
        take, tot = sample_prefix(src_size, int(args.cpu_sample_gib * (1 << 30)))
        host = np.concatenate([d_src[int(src_off[i]):int(src_off[i]) + int(src_size[i])].cpu().numpy() for i in range(take)])
        r = run_cpu_bench(host, src_size[:take].astype(np.int64), opts, reps=3)
        if r:
            cpu = {"value": combined_gbs(r["raw_bytes"], r["encode_s"], r["decode_s"]), "unit": "GB/s", "cores": r["threads"], "kind": "reference",
                   "encode_gbs": r["encode_gbs"], "decode_gbs": r["decode_gbs"], "ratio": r["ratio"],
                   "sample": "first %d reads (%.2f GiB raw) of this run's batch, copied to the host; oracle/_ref (unmodified reference, -O3 -mssse3, libzstd 1.5.5), one vbz_*_sized call per read, best of 3" % (take, r["raw_bytes"] / 2**30)}
            # every GPU-written frame of the sample must decode with the reference to the original bytes
            ctake = min(take, 400)
            gpu_chunks = np.concatenate([d_cmp[int(dst_off[i]):int(dst_off[i]) + int(csize[0][i])].cpu().numpy() for i in range(ctake)])
            ctot = int(src_size[:ctake].astype(np.int64).sum())
            x = run_cpu_decode(gpu_chunks, csize[0][:ctake], host[:ctot], src_size[:ctake].astype(np.int64), opts, reps=1)
            cpu["cross_decode_ok"] = bool(x is not None and x["errors"] == 0)
            cpu["cross_decode_sample"] = "%d GPU-written chunks decoded by the reference's vbz_decompress_sized and compared with the input" % ctake
            if x:
                cpu["reference_decode_of_gpu_frames_gbs"] = x["decode_gbs"]

    dtype = {"uint32": "uint32", "int8": "int8"}.get(args.workload, "int16")
    line = {
        "metric": "raw %s GB/s encode+decode" % dtype, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": dtype, "data": "synthetic",
        "config": workload_config(args, opts),
        "encode_gbs": all_raw / (enc_m * 1e-3) / 1e9, "decode_gbs": all_raw / (dec_m * 1e-3) / 1e9, "ratio": all_comp / all_raw,
        "raw_bytes_per_gpu": raw_bytes, "roofline": roofline, "transform_stage": transform_stage, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks,
        "gpu_launches": int(all_launches), "host_affinity": numa,
    }
    emit_line(json.dumps(line))
    D.close()


def run_e2e(args, D, batch, opts, d_src, src_off, src_size, n):
    """Same metric through vbz_b200_compress_batch / vbz_b200_decompress_batch with HOST buffers, on every rank at once."""
    import torch
    take, tot = sample_prefix(src_size, int(args.e2e_sample_gib * (1 << 30)))
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
    D.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        step()
    dt = (time.perf_counter() - t0) / reps
    (dt_max,) = D.reduce([dt], "max")
    all_tot, all_comp, all_take = D.reduce([tot, comp, take], "sum")
    return {"value": 2.0 * all_tot / dt_max / 1e9, "unit": "GB/s", "h2d_bytes_per_step": int(all_tot + all_comp), "d2h_bytes_per_step": int(all_comp + all_tot),
            "ms_per_step": dt_max * 1e3, "ranks": D.world, "per_rank_gbs_this_rank": 2.0 * tot / dt / 1e9,
            "sample": "on each of the %d ranks at the same time: first %d reads (%.2f GiB raw) of the rank's batch as pinned HOST buffers through "
                      "vbz_b200_compress_batch + vbz_b200_decompress_batch (wall clock, %d reps; bytes summed over ranks, time = max over ranks)"
                      % (D.world, take, tot / 2**30, reps)}


# --------------------------------------------------------------------------------------------------
def main_h5z(args):
    """BASELINE configs[4]: a decode-heavy job of reference-written H5Z chunk buffers sharded by read index."""
    import torch
    from vbz_compression_b200 import Batch, load
    from vbz_compression_b200.shard import balanced_ranges

    D = Dist()
    world, rank, local, dev = D.world, D.rank, D.local, D.dev
    numa = pin_to_gpu_numa(local)
    lib = load()
    batch = Batch(local, lib=lib)
    opts = options_for("h5z-decode")
    cd_variants = [(0, 2, 1, 1), (1, 2, 1, 1)]

    # ---- the pool of distinct chunks, written by the reference (rank 0 makes it, everyone reads it) -----------------------
    pool_dir = os.path.join(tempfile.gettempdir(), "vbz_b200_pool_%s" % os.environ.get("MASTER_PORT", str(os.getpid())))
    if rank == 0:
        pool = reference_pool(args, opts, SEED)
        assert pool is not None, "oracle/_ref/cpu_bench is needed to write the reference chunks"
        os.makedirs(pool_dir, exist_ok=True)
        for name, arr in zip(("raw.u8", "lens.i64", "chunks.u8", "csz.u32"), pool):
            arr.tofile(os.path.join(pool_dir, name))
    D.barrier()
    raw = np.fromfile(os.path.join(pool_dir, "raw.u8"), np.uint8)
    lens_b = np.fromfile(os.path.join(pool_dir, "lens.i64"), np.int64)
    chunks = np.fromfile(os.path.join(pool_dir, "chunks.u8"), np.uint8)
    csz = np.fromfile(os.path.join(pool_dir, "csz.u32"), np.uint32)
    P = len(lens_b)
    raw_off = np.concatenate([[0], np.cumsum(lens_b)[:-1]]).astype(np.int64)
    c_off = np.concatenate([[0], np.cumsum(csz.astype(np.int64))[:-1]]).astype(np.int64)

    # ---- the job: read i = pool chunk i mod P; this rank's contiguous, byte-balanced range ---------------------------------
    job_sizes = np.tile(lens_b, (args.reads + P - 1) // P)[:args.reads]
    lo, hi = balanced_ranges(job_sizes, world)[rank]
    calls = []  # pool index ranges, one decode call each
    i = lo
    while i < hi:
        a = i % P
        b = min(P, a + (hi - i))
        calls.append((a, b))
        i += b - a
    shard_raw = int(job_sizes[lo:hi].sum())
    shard_comp = int(sum(int(csz[a:b].astype(np.int64).sum()) for a, b in calls))

    # ---- resident in HBM: the pool's chunks (4-byte aligned slots), an output buffer for one pass over the pool -----------
    a16 = lambda v: (v + 15) & ~np.uint64(15)
    d_coff = np.concatenate([[0], np.cumsum(a16(csz.astype(np.uint64)))[:-1]]).astype(np.uint64)
    d_ooff = np.concatenate([[0], np.cumsum(a16(lens_b.astype(np.uint64)))[:-1]]).astype(np.uint64)
    d_chunks = torch.zeros(int(a16(csz.astype(np.uint64)).sum()) + 16, dtype=torch.uint8, device=dev)
    h_chunks = torch.from_numpy(chunks)
    for k in range(P):
        d_chunks[int(d_coff[k]):int(d_coff[k]) + int(csz[k])].copy_(h_chunks[int(c_off[k]):int(c_off[k]) + int(csz[k])])
    d_out = torch.empty(int(a16(lens_b.astype(np.uint64)).sum()) + 16, dtype=torch.uint8, device=dev)
    d_res = torch.zeros(P, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    out_sz = lens_b.astype(np.uint32)
    from vbz_compression_b200 import CompressionOptions

    def cd_opts(cd):
        return CompressionOptions(bool(cd[2]), cd[1], cd[3], cd[0])

    def decode_range(a, b, cd):
        batch.decompress_device(d_chunks.data_ptr(), d_coff[a:b], csz[a:b], d_out.data_ptr(), d_ooff[a:b], out_sz[a:b], d_res[a:].data_ptr(), cd_opts(cd), True, stream)

    # correctness: one pass over the pool, every byte compared with what the reference was given
    d_raw = torch.from_numpy(raw).to(dev)
    for cd in cd_variants:
        d_out.zero_()
        decode_range(0, P, cd)
        torch.cuda.synchronize()
        res = d_res.cpu().numpy().view(np.uint32)
        assert (res == out_sz).all(), "decode reported errors"
        if not args.no_verify:
            for k in range(P):
                assert torch.equal(d_out[int(d_ooff[k]):int(d_ooff[k]) + int(lens_b[k])], d_raw[int(raw_off[k]):int(raw_off[k]) + int(lens_b[k])]), "chunk %d differs" % k
    del d_raw

    def step(k):
        for a, b in calls:
            decode_range(a, b, cd_variants[k % 2])

    for k in range(args.warmup):
        step(k)
    torch.cuda.synchronize()
    D.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start_and_wait()
    D.barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    l0 = batch.kernel_launches
    torch.cuda.synchronize()
    ev[0].record()
    for k in range(args.steps):
        step(k); ev[k + 1].record()
    torch.cuda.synchronize()
    launches = batch.kernel_launches - l0
    clocks = sampler.stop() if rank == 0 else None
    total_ms = ev[0].elapsed_time(ev[-1])
    (total_ms,) = D.reduce([total_ms], "max")
    all_raw, all_comp, all_launches = D.reduce([shard_raw, shard_comp, launches], "sum")
    value = all_raw * args.steps / (total_ms * 1e-3) / 1e9

    # stage split of one full pass over the pool (this rank)
    f1, f2 = ctypes.c_float(), ctypes.c_float()
    lib.vbz_b200_set_stage_timing(batch.ctx, 1)
    s1_ms, z_ms = [], []
    for k in range(3):
        decode_range(0, P, cd_variants[k % 2])
        assert lib.vbz_b200_stage_times(batch.ctx, ctypes.byref(f1), ctypes.byref(f2)) == 0
        s1_ms.append(f1.value); z_ms.append(f2.value)
    lib.vbz_b200_set_stage_timing(batch.ctx, 0)
    pool_raw, pool_comp = int(lens_b.sum()), int(csz.astype(np.int64).sum())

    # ---- end to end: pinned host chunk buffers -> pinned host samples, through the batch form of the filter callback ------
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e_h5z(args, D, batch, chunks, csz, c_off, lens_b, raw, raw_off, calls, cd_variants)
    if rank != 0:
        D.close()
        return

    peak, peak_src = measured_peak()
    # stage-1 bytes of the pool (control bytes + data bytes) = the content size each frame announces in its header
    s1_pool = int(sum(frame_content_size(chunks[int(c_off[k]) + 4:int(c_off[k]) + 24]) for k in range(P)))
    groups = {"zstd decode (walk + units + general kernels)": (s1_pool + pool_comp, float(np.mean(z_ms))),
              "svb16_decode_kernel": (pool_raw + s1_pool, float(np.mean(s1_ms)))}
    roofline = stage_roofline(groups, peak, peak_src, "one pass over the pool (%d chunks, %.2f GiB raw) on this rank" % (P, pool_raw / 2**30))

    cpu = None
    if not args.no_cpu_baseline:
        take, tot = sample_prefix(lens_b, int(args.cpu_sample_gib * (1 << 30)))
        ctake = int(csz[:take].astype(np.int64).sum())
        r = run_cpu_decode(chunks[:ctake], csz[:take], raw[:tot], lens_b[:take], opts, reps=3)
        if r:
            cpu = {"value": r["decode_gbs"], "unit": "GB/s", "cores": r["threads"], "kind": "reference", "decode_gbs": r["decode_gbs"], "ratio": r["ratio"],
                   "sample": "first %d chunks (%.2f GiB raw) of the pool; oracle/_ref, one vbz_decompress_sized call per chunk, all host cores, best of 3" % (take, tot / 2**30)}

    line = {
        "metric": "raw int16 GB/s decode", "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int16", "data": "synthetic",
        "config": workload_config(args, opts), "decode_gbs": value, "ratio": pool_comp / pool_raw,
        "job": {"reads": args.reads, "raw_bytes": int(all_raw), "pool_chunks": P, "this_rank": {"first": int(lo), "last": int(hi), "raw_bytes": shard_raw, "calls": len(calls)}},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks, "gpu_launches": int(all_launches), "host_affinity": numa,
    }
    emit_line(json.dumps(line))
    D.close()


def run_e2e_h5z(args, D, batch, chunks, csz, c_off, lens_b, raw, raw_off, calls, cd_variants):
    import torch
    lib = batch.lib
    P = len(lens_b)
    h_chunks = torch.from_numpy(chunks.copy()).pin_memory()
    h_out = torch.empty(int(lens_b.sum()), dtype=torch.uint8).pin_memory()
    vp, u32 = ctypes.c_void_p, ctypes.c_uint32
    srcp = (vp * P)(*[h_chunks.data_ptr() + int(o) for o in c_off])
    dstp = (vp * P)(*[h_out.data_ptr() + int(o) for o in raw_off])
    ssz = (u32 * P)(*[int(x) for x in csz])
    dcap = (u32 * P)(*[int(x) for x in lens_b])
    res = (u32 * P)()
    H5Z_FLAG_REVERSE = 0x0100

    def off(arr, a, typ):
        return ctypes.cast(ctypes.byref(arr, a * ctypes.sizeof(typ)), ctypes.POINTER(typ))

    def call(a, b, cd):
        cdv = (ctypes.c_uint * 4)(*cd)
        rc = lib.vbz_b200_h5z_filter_batch(batch.ctx, H5Z_FLAG_REVERSE, 4, cdv, b - a, off(srcp, a, vp), off(ssz, a, u32), off(dstp, a, vp), off(dcap, a, u32), off(res, a, u32))
        assert rc == 0, lib.vbz_b200_last_error()

    call(0, P, cd_variants[0])
    assert (np.frombuffer(res, np.uint32) == lens_b.astype(np.uint32)).all(), "e2e decode reported errors"
    assert torch.equal(h_out, torch.from_numpy(raw)), "e2e decode mismatch"
    # the rank's whole range if that takes seconds, else as many of its calls as fit ~10 s (measured on the first call)
    t0 = time.perf_counter()
    call(calls[0][0], calls[0][1], cd_variants[1])
    first = time.perf_counter() - t0
    first_bytes = int(lens_b[calls[0][0]:calls[0][1]].sum())
    budget_calls = len(calls)
    if first > 0 and first_bytes > 0:
        rate = first_bytes / first
        done, k = 0, 0
        while k < len(calls) and done / rate < 10.0:
            done += int(lens_b[calls[k][0]:calls[k][1]].sum()); k += 1
        budget_calls = max(1, k)
    (bc,) = D.reduce([-budget_calls], "max")  # the same number of calls on every rank (the smallest budget)
    budget_calls = min(len(calls), int(-bc))
    D.barrier()
    t0 = time.perf_counter()
    rawb = compb = 0
    for k in range(budget_calls):
        a, b = calls[k]
        call(a, b, cd_variants[k % 2])
        rawb += int(lens_b[a:b].sum()); compb += int(csz[a:b].astype(np.int64).sum())
    dt = time.perf_counter() - t0
    (dt_max,) = D.reduce([dt], "max")
    all_raw, all_comp = D.reduce([rawb, compb], "sum")
    return {"value": all_raw / dt_max / 1e9, "unit": "GB/s", "h2d_bytes_per_step": int(all_comp), "d2h_bytes_per_step": int(all_raw), "ms_per_step": dt_max * 1e3,
            "ranks": D.world, "per_rank_gbs_this_rank": rawb / dt / 1e9,
            "sample": "on each of the %d ranks at the same time: the first %d of the rank's %d decode calls (one pass over the chunk pool each at most), pinned HOST chunk "
                      "buffers in, pinned HOST samples out, through vbz_b200_h5z_filter_batch (flags = H5Z_FLAG_REVERSE, cd_values alternating (0,2,1,1) / (1,2,1,1)); "
                      "wall clock, bytes summed over ranks, time = max over ranks" % (D.world, budget_calls, len(calls))}


_JSON_OUT = None


def emit_line(text):
    """The one JSON line goes to the process's ORIGINAL stdout (see the guard in __main__)."""
    out = _JSON_OUT or sys.stdout
    out.write(text + "\n")
    out.flush()


if __name__ == "__main__":
    # stdout carries the JSON line and nothing else: file descriptor 1 is pointed at stderr for the duration of the run, so that whatever a
    # library prints there (NCCL's version banner does not honour NCCL_DEBUG_FILE in every build) cannot end up in front of the line
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    a = default_reads(parse_args())
    if a.impl == "reference":
        main_reference(a)
    elif a.workload == "h5z-decode":
        main_h5z(a)
    else:
        main_codec(a)
