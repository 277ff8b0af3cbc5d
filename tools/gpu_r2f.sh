#!/bin/bash
# ncu --set full of the new zstd kernels (one launch each), small batch
set -u
mkdir -p gpurun_out
T=${1:-r2f}
ncu --set full --clock-control none --import-source on -k regex:'zstd_enc_match_kernel|zstd_enc_seq_kernel|zstd_seq_kernel|zstd_seqexec_kernel' --launch-skip 4 --launch-count 4 -f -o gpurun_out/${T}_zseq \
    python bench.py --workload full --reads 6000 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-verify > gpurun_out/${T}_ncu_zseq.log 2>&1
ls -la gpurun_out/${T}_zseq.ncu-rep; tail -2 gpurun_out/${T}_ncu_zseq.log | cut -c1-300
