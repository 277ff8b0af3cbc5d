#!/bin/bash
# ncu --set full of the decode kernels on reference-written frames (h5z-decode workload, small pool)
set -u
mkdir -p gpurun_out
T=${1:-h5zprof}; K=${2:-zstd_seq_kernel}
ncu --set full --clock-control none --import-source on -k regex:"$K" --launch-skip 2 --launch-count 3 -f -o gpurun_out/${T}_k \
    python bench.py --workload h5z-decode --pool-gib 1.0 --reads 20000 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-verify > gpurun_out/${T}_ncu.log 2>&1
ls -la gpurun_out/${T}_k.ncu-rep; tail -3 gpurun_out/${T}_ncu.log | cut -c1-400
