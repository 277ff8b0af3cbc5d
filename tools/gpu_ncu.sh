#!/bin/bash
# ncu passes (never a bench value): launch list of the bench command, and --set full captures of the codec kernels
set -x
mkdir -p gpurun_out
R=${1:-20000}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'svb|zstd|vbz' -c 400 --csv --log-file gpurun_out/launches_transform.csv \
    python bench.py --reads $R --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_b1.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'svb|zstd|vbz' -c 400 --csv --log-file gpurun_out/launches_full.csv \
    python bench.py --workload full --reads $R --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_b2.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'svb16_(en|de)code_kernel' -s 2 -c 2 -o gpurun_out/prof_svb16 -f \
    python bench.py --reads $R --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_b3.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'zstd_(en|de)code_kernel' -s 2 -c 2 -o gpurun_out/prof_zstd -f \
    python bench.py --workload full --reads $R --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_b4.log 2>&1
ls -la gpurun_out
