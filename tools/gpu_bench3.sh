#!/bin/bash
# the three bench lines as the driver runs them (full default, then the other two workloads) + the reference arm
set -u
mkdir -p gpurun_out
T=${1:-b3}
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/${T}_bench_full.json 2> gpurun_out/${T}_bench_full.err; echo "full rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/${T}_bench_full.json'))
print('value',round(d['value'],1),'enc',round(d.get('encode_gbs',0),1),'dec',round(d['decode_gbs'],1),'ratio',round(d['ratio'],5),'e2e',round(d['e2e']['value'],2), 'cpu', d.get('cpu_baseline',{}) and d['cpu_baseline'].get('value'))
for k,v in d['roofline']['stages'].items(): print('   ',k,round(v['ms'],2),'ms',round(v['frac'],3))
PY
