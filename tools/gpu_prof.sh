#!/bin/bash
# quick profile visit: stage times of a small resident batch + ncu --set full of chosen kernels
# usage: gpu_prof.sh TAG 'kernel-regex' [reads]
set -u
mkdir -p gpurun_out
T=${1:-prof}; K=${2:-zstd_enc_match_kernel}; R=${3:-6000}
python bench.py --workload full --reads $R --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-verify > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/${T}_bench.json'))
print('value',round(d['value'],1),'enc',round(d.get('encode_gbs',0),1),'dec',round(d['decode_gbs'],1),'ratio',round(d['ratio'],5))
for k,v in d['roofline']['stages'].items(): print('   ',k,round(v['ms'],2),'ms',round(v['frac'],3))
PY
ncu --set full --clock-control none --import-source on -k regex:"$K" --launch-skip 4 --launch-count 4 -f -o gpurun_out/${T}_k \
    python bench.py --workload full --reads $R --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-verify > gpurun_out/${T}_ncu.log 2>&1
ls -la gpurun_out/${T}_k.ncu-rep
