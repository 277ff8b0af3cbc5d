#!/bin/bash
# quick visit: parity tests + the full-VBZ bench line (+ optional launch list)
set -u
mkdir -p gpurun_out
T=${1:-rXX}
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/${T}_gpu_tests.log
tail -3 gpurun_out/${T}_gpu_tests.log
python bench.py --workload full --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_full.json 2> gpurun_out/${T}_bench_full.err
cat gpurun_out/${T}_bench_full.json; tail -5 gpurun_out/${T}_bench_full.err
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'svb|zstd|vbz' -c 400 --csv --log-file gpurun_out/${T}_launches_full.csv \
      python bench.py --workload full --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${T}_ncu_bench_full.log 2>&1
python tools/ncu_summary.py launches gpurun_out/${T}_launches_full.csv gpurun_out/${T}_launches_full.txt; cat gpurun_out/${T}_launches_full.txt
