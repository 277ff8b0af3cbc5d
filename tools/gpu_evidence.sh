#!/bin/bash
# One visit that produces what profiles/ holds for a round: GPU parity tests, launch lists of the three workloads, an ncu --set full
# capture of every kernel of the full path (one launch each), then the bench lines (NOT under ncu) and the reference arm.
# usage: gpu_evidence.sh TAG
set -u
mkdir -p gpurun_out
T=${1:-r12}
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/${T}_pytest_gpu.log; cat gpurun_out/${T}_pytest_gpu.log
for W in full h5z-decode uint32; do
  EXTRA=""; [ "$W" = h5z-decode ] && EXTRA="--pool-gib 1.0"
  R=20000; [ "$W" = uint32 ] && R=8000
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'svb|zstd|vbz' -c 500 --csv --log-file gpurun_out/${T}_launches_${W}.csv \
     python bench.py --workload $W --reads $R $EXTRA --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-verify > gpurun_out/${T}_ncu_${W}.log 2>&1
  python tools/ncu_summary.py launches gpurun_out/${T}_launches_${W}.csv gpurun_out/${T}_launches_${W}.txt
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'svb16_encode_kernel|svb16_decode_kernel|zstd_' --launch-skip 11 --launch-count 11 -f -o gpurun_out/${T}_full \
    python bench.py --workload full --reads 6000 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-verify > gpurun_out/${T}_ncu_setfull.log 2>&1
python tools/ncu_summary.py raw gpurun_out/${T}_full.ncu-rep gpurun_out/${T}_ncu_kernels.txt
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/${T}_bench_full.json 2> gpurun_out/${T}_bench_full.err; echo "full rc=$?"
timeout 1200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_ref_full.json 2> gpurun_out/${T}_bench_ref_full.err; echo "ref rc=$?"
timeout 1200 python bench.py --workload h5z-decode --pool-gib 8 --steps 3 --warmup 2 > gpurun_out/${T}_bench_h5z.json 2> gpurun_out/${T}_bench_h5z.err; echo "h5z rc=$?"
timeout 1200 python bench.py --workload uint32 --steps 3 --warmup 2 > gpurun_out/${T}_bench_uint32.json 2> gpurun_out/${T}_bench_uint32.err; echo "uint32 rc=$?"
python - <<PY
import json
for w in ("full", "ref_full", "h5z", "uint32"):
    try:
        d = json.load(open("gpurun_out/${T}_bench_%s.json" % w))
        print(w, "value", round(d["value"], 1), "ratio", d.get("ratio"), "e2e", round(d["e2e"]["value"], 2) if d.get("e2e") else None, "cpu", d.get("cpu_baseline") and d["cpu_baseline"].get("value"),
              "frac", d.get("roofline") and round(d["roofline"]["frac"], 3))
    except Exception as e:
        print(w, "failed", e)
PY
