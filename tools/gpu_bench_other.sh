#!/bin/bash
set -u
mkdir -p gpurun_out
T=${1:-bo}
for W in h5z-decode uint32; do
timeout 1200 python bench.py --workload $W --steps 3 --warmup 2 > gpurun_out/${T}_bench_$W.json 2> gpurun_out/${T}_bench_$W.err; echo "$W rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/${T}_bench_$W.json'))
print('$W value',round(d['value'],1),'dec',round(d['decode_gbs'],1),'ratio',round(d['ratio'],5),'e2e',round(d['e2e']['value'],2), 'cpu', d.get('cpu_baseline') and d['cpu_baseline'].get('value'))
for k,v in d['roofline']['stages'].items(): print('   ',k,round(v['ms'],2),'ms',round(v['frac'],3))
PY
done
