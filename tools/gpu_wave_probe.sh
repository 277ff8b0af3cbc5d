#!/bin/bash
# h5z-decode and full e2e under a few wave sizing settings (VBZ_B200_WAVE_CAP_MIB / _DIV / _FIRST_MIB): usage gpu_wave_probe.sh "cap:div:first" ...
for S in "$@"; do
  IFS=: read C D F <<< "$S"
  export VBZ_B200_WAVE_CAP_MIB=$C VBZ_B200_WAVE_DIV=$D VBZ_B200_WAVE_FIRST_MIB=$F
  timeout 600 python bench.py --workload h5z-decode --pool-gib 4 --steps 1 --warmup 1 --no-cpu-baseline --no-verify 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('h5z', '$S', 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],2))"
done
