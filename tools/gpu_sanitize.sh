#!/bin/bash
# compute-sanitizer over the GPU tests of the zstd stage, the host batches and the device-resident waves: memcheck, then racecheck
set -u
mkdir -p gpurun_out
T=${1:-san}
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_zstd_decode.py tests/test_zstd_encode.py tests/test_host_batch.py tests/test_fast5.py tests/test_device_api.py -m gpu -x -q > gpurun_out/${T}_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/${T}_memcheck.log; tail -4 gpurun_out/${T}_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 --kernel-name kns=zstd python -m pytest tests/test_zstd_decode.py tests/test_zstd_encode.py -m gpu -x -q > gpurun_out/${T}_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/${T}_racecheck.log; tail -4 gpurun_out/${T}_racecheck.log
