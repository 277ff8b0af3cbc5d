#!/bin/bash
# launch list (ncu gpu__time_duration) of one bench workload: gpu_launches.sh TAG [workload] [reads]
set -u
mkdir -p gpurun_out
T=${1:-ll}; W=${2:-full}; R=${3:-20000}
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'svb|zstd|vbz' -c 400 --csv --log-file gpurun_out/${T}_launches_${W}.csv \
   python bench.py --workload $W --reads $R --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-verify > gpurun_out/${T}_ncu_${W}.log 2>&1
python tools/ncu_summary.py launches gpurun_out/${T}_launches_${W}.csv gpurun_out/${T}_launches_${W}.txt; head -20 gpurun_out/${T}_launches_${W}.txt
