#!/bin/bash
# last visit of a round: the GPU test suite, the default bench line and the reference-written chunks, all of the final code.  usage: gpu_final.sh TAG
set -u
mkdir -p gpurun_out
T=${1:-final}
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/${T}_pytest_gpu.log; cat gpurun_out/${T}_pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${T}_bench_full.json 2> gpurun_out/${T}_bench_full.err; echo "full rc=$?"
timeout 1200 python bench.py --workload h5z-decode --pool-gib 8 --steps 3 --warmup 2 > gpurun_out/${T}_bench_h5z.json 2> gpurun_out/${T}_bench_h5z.err; echo "h5z rc=$?"
python - <<PY
import json
for w in ("full", "h5z"):
    try:
        d = json.load(open("gpurun_out/${T}_bench_%s.json" % w))
        print(w, "value", round(d["value"], 1), "ratio", d.get("ratio"), "e2e", round(d["e2e"]["value"], 2) if d.get("e2e") else None, "cpu", d.get("cpu_baseline") and d["cpu_baseline"].get("value"),
              {k: round(v["ms"], 1) for k, v in d["roofline"]["stages"].items()})
    except Exception as e:
        print(w, "failed", e)
PY
