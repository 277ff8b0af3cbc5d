#!/bin/bash
# quick visit: gpu tests of the zstd stage + launch list of the full workload + one bench line
set -u
mkdir -p gpurun_out
T=${1:-r2g}
timeout 1200 python -m pytest tests/test_zstd_encode.py tests/test_zstd_decode.py tests/test_api.py -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -3 gpurun_out/${T}_pytest.log
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'svb|zstd|vbz' -c 400 --csv --log-file gpurun_out/${T}_launches_full.csv \
   python bench.py --workload full --reads 20000 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-verify > gpurun_out/${T}_ncu_full.log 2>&1
python tools/ncu_summary.py launches gpurun_out/${T}_launches_full.csv gpurun_out/${T}_launches_full.txt; head -16 gpurun_out/${T}_launches_full.txt
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_full.json 2> gpurun_out/${T}_bench_full.err; echo "full rc=$?"; tail -5 gpurun_out/${T}_bench_full.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/%s_bench_full.json' % "${T}"))
print('value',round(d['value'],1),'enc',round(d.get('encode_gbs',0),1),'dec',round(d['decode_gbs'],1),'ratio',round(d['ratio'],5),'e2e',round(d['e2e']['value'],2))
for k,v in d['roofline']['stages'].items(): print('   ',k,round(v['ms'],2),'ms',round(v['frac'],3))
PY
