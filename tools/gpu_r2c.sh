#!/bin/bash
# per-kernel times of the h5z-decode workload (reference-written frames)
set -u
mkdir -p gpurun_out
T=${1:-r2c}
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'svb|zstd|vbz' -c 300 --csv --log-file gpurun_out/${T}_launches_h5z.csv \
   python bench.py --workload h5z-decode --reads 5000 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-verify > gpurun_out/${T}_ncu_h5z.log 2>&1
python tools/ncu_summary.py launches gpurun_out/${T}_launches_h5z.csv gpurun_out/${T}_launches_h5z.txt; head -20 gpurun_out/${T}_launches_h5z.txt; tail -3 gpurun_out/${T}_ncu_h5z.log
