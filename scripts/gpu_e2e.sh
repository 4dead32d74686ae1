#!/bin/bash
# usage: bash scripts/gpu_e2e.sh <tag> chunk...   -- full bench (e2e + side measurements) per e2e chunk size
mkdir -p gpurun_out
TAG=$1; shift
for CH in "$@"; do
  timeout 300 python bench.py --steps 20 --warmup 3 --e2e-chunk $CH > gpurun_out/e2e_${TAG}_$CH.json 2> gpurun_out/e2e_${TAG}_$CH.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/e2e_${TAG}_$CH.json').read().strip().splitlines()[-1])
print('chunk', $CH, 'value', round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'e2e ms', round(d['e2e']['ms_per_step'],2), 'stages', {k: round(v['median']*1e3,1) for k,v in d['stages_ms'].items()})
print('lidar', d.get('lidar'))
PY
  tail -2 gpurun_out/e2e_${TAG}_$CH.err
done
