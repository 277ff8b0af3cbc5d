#!/bin/bash
# round 2, visit A: parity tests, smoke, the default bench line at 100k reads, and first lines of the new workloads
set -u
mkdir -p gpurun_out
T=${1:-r2a}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -4 gpurun_out/${T}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${T}_smoke.log; tail -2 gpurun_out/${T}_smoke.log
nvidia-smi --query-gpu=memory.total,memory.used --format=csv > gpurun_out/${T}_mem.log; nproc >> gpurun_out/${T}_mem.log; free -g >> gpurun_out/${T}_mem.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${T}_bench_full.json 2> gpurun_out/${T}_bench_full.err; echo "full rc=$?"; tail -c 1500 gpurun_out/${T}_bench_full.json; tail -5 gpurun_out/${T}_bench_full.err
timeout 900 python bench.py --workload uint32 --steps 5 --warmup 3 > gpurun_out/${T}_bench_uint32.json 2> gpurun_out/${T}_bench_uint32.err; echo "uint32 rc=$?"; tail -c 1500 gpurun_out/${T}_bench_uint32.json; tail -5 gpurun_out/${T}_bench_uint32.err
timeout 900 python bench.py --workload h5z-decode --reads 50000 --steps 3 --warmup 1 > gpurun_out/${T}_bench_h5z.json 2> gpurun_out/${T}_bench_h5z.err; echo "h5z rc=$?"; tail -c 1500 gpurun_out/${T}_bench_h5z.json; tail -5 gpurun_out/${T}_bench_h5z.err
