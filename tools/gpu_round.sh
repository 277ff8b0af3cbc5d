#!/bin/bash
# One GPU-box pass: parity tests, bench (both workloads), ncu launch list + full captures. Outputs -> gpurun_out/
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt
nproc >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_transform.json 2> gpurun_out/bench_transform.err; tail -c 3000 gpurun_out/bench_transform.json
timeout 600 python bench.py --workload full --steps 3 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -c 3000 gpurun_out/bench_full.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_transform.json 2>&1
timeout 600 python bench.py --impl reference --workload full --steps 2 --warmup 1 > gpurun_out/bench_ref_full.json 2>&1
