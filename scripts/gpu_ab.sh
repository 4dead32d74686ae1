#!/bin/bash
# usage: bash scripts/gpu_ab.sh <tag> "<ENV...>" ["<ENV...>" ...]   -- GPU tests once, then one short bench per env setting
mkdir -p gpurun_out
TAG=$1; shift
timeout 900 python -m pytest tests -m gpu -x -q --timeout 180 2>&1 | tail -8
i=0
for E in "$@"; do
  echo "=== [$i] $E"
  env $E timeout 200 python bench.py --steps 20 --warmup 3 --no-extras 2> gpurun_out/ab_${TAG}_$i.err | tee gpurun_out/ab_${TAG}_$i.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(round(d['value']), 'frames/s', round(d['ms_per_step']*1e3,1), 'us/step', {k: round(v['median']*1e3,1) for k,v in d['stages_ms'].items()})"
  tail -2 gpurun_out/ab_${TAG}_$i.err
  i=$((i+1))
done
