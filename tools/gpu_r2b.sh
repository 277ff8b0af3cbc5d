#!/bin/bash
# round 2, visit B: parity tests + memcheck of the decode kernels + h5z-decode bench
set -u
mkdir -p gpurun_out
T=${1:-r2b}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -4 gpurun_out/${T}_pytest.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_zstd_decode.py -m gpu -x -q > gpurun_out/${T}_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/${T}_memcheck.log; tail -5 gpurun_out/${T}_memcheck.log
timeout 900 python bench.py --workload h5z-decode --reads 100000 --steps 5 --warmup 2 > gpurun_out/${T}_bench_h5z.json 2> gpurun_out/${T}_bench_h5z.err; echo "h5z rc=$?"; tail -c 2500 gpurun_out/${T}_bench_h5z.json; tail -5 gpurun_out/${T}_bench_h5z.err
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_full.json 2> gpurun_out/${T}_bench_full.err; echo "full rc=$?"; tail -c 1200 gpurun_out/${T}_bench_full.json; tail -5 gpurun_out/${T}_bench_full.err
