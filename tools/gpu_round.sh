#!/bin/bash
# One GPU-box visit: parity tests, ncu launch lists of the bench command, then the two bench lines (not under ncu).
set -u
mkdir -p gpurun_out
T=${1:-r02}
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/${T}_gpu_tests.log
for W in transform full; do
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'svb|zstd|vbz' -c 400 --csv --log-file gpurun_out/${T}_launches_${W}.csv \
      python bench.py --workload $W --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${T}_ncu_bench_${W}.log 2>&1
done
python bench.py --steps 5 --warmup 3 > gpurun_out/${T}_bench_transform.json 2> gpurun_out/${T}_bench_transform.err
python bench.py --workload full --steps 3 --warmup 3 > gpurun_out/${T}_bench_full.json 2> gpurun_out/${T}_bench_full.err
tail -3 gpurun_out/${T}_gpu_tests.log
cat gpurun_out/${T}_bench_transform.json gpurun_out/${T}_bench_full.json
