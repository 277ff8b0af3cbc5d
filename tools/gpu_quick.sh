#!/bin/bash
# quick GPU pass: parity tests + the full-VBZ bench (and optionally the transform bench with $1 = all)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --workload full --steps 3 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -c 2500 gpurun_out/bench_full.json; tail -5 gpurun_out/bench_full.err
if [ "$1" = "all" ]; then
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_transform.json 2> gpurun_out/bench_transform.err; tail -c 2500 gpurun_out/bench_transform.json; tail -5 gpurun_out/bench_transform.err
fi
