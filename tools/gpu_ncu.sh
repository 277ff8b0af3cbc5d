#!/bin/bash
# ncu --set full captures of the dominant kernels (one launch each, after warm-up launches), small batches.
# usage: gpu_ncu.sh TAG [zstd|svb16|all]
set -u
mkdir -p gpurun_out
T=${1:-r02}
W=${2:-all}
if [ "$W" = all ] || [ "$W" = zstd ]; then
ncu --set full --clock-control none --import-source on -k regex:'zstd_units_kernel|zstd_encode_kernel' --launch-skip 2 --launch-count 2 -f -o gpurun_out/${T}_zstd \
    python bench.py --workload full --reads 8000 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${T}_ncu_zstd.log 2>&1
fi
if [ "$W" = all ] || [ "$W" = svb16 ]; then
ncu --set full --clock-control none --import-source on -k regex:'svb16_encode_kernel|svb16_decode_kernel|svb16_tile_sums_kernel' --launch-skip 3 --launch-count 3 -f -o gpurun_out/${T}_svb16 \
    python bench.py --reads 20000 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${T}_ncu_svb16.log 2>&1
fi
ls -la gpurun_out
