#!/bin/bash
# round 2, visit D: parity tests, memcheck + racecheck of the zstd kernels, launch lists and bench lines of full and h5z-decode
set -u
mkdir -p gpurun_out
T=${1:-r2d}
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -4 gpurun_out/${T}_pytest.log
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_zstd_decode.py tests/test_zstd_encode.py -m gpu -x -q > gpurun_out/${T}_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/${T}_memcheck.log; tail -4 gpurun_out/${T}_memcheck.log
timeout 600 python tools/ratio_probe.py gpu > gpurun_out/${T}_ratio.log 2>&1; tail -18 gpurun_out/${T}_ratio.log
for W in full h5z-decode; do
  R=20000; [ $W = h5z-decode ] && R=5000
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'svb|zstd|vbz' -c 400 --csv --log-file gpurun_out/${T}_launches_${W}.csv \
     python bench.py --workload $W --reads $R --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-verify > gpurun_out/${T}_ncu_${W}.log 2>&1
  python tools/ncu_summary.py launches gpurun_out/${T}_launches_${W}.csv gpurun_out/${T}_launches_${W}.txt; head -24 gpurun_out/${T}_launches_${W}.txt
done
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${T}_bench_full.json 2> gpurun_out/${T}_bench_full.err; echo "full rc=$?"; tail -c 600 gpurun_out/${T}_bench_full.json; tail -5 gpurun_out/${T}_bench_full.err
timeout 900 python bench.py --workload h5z-decode --reads 100000 --steps 5 --warmup 2 > gpurun_out/${T}_bench_h5z.json 2> gpurun_out/${T}_bench_h5z.err; echo "h5z rc=$?"; tail -c 600 gpurun_out/${T}_bench_h5z.json; tail -5 gpurun_out/${T}_bench_h5z.err
