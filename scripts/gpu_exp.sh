#!/bin/bash
# usage: bash scripts/gpu_exp.sh "ENV1=a ENV2=b" "ENV1=c" ...   -- kernel-stage timings of bench.py per env setting
# (tight timeouts everywhere: a deadlocked kernel must not eat the GPU budget)
mkdir -p gpurun_out
timeout 60 python -u scripts/debug_steps.py all 2 2>&1 | tail -6 || { echo "debug_steps failed/hung"; exit 1; }
timeout 240 python -m pytest tests/test_gpu_voxel_pool.py -x -q --timeout 40 2>&1 | tail -6
if [ "${PIPESTATUS[0]}" != "0" ]; then echo "TESTS FAILED - skipping benches"; exit 1; fi
for CFG in "$@"; do
  echo "== $CFG"
  env $CFG timeout 90 python bench.py --steps 20 --warmup 3 --no-extras 2>&1 | python -c "
import json,sys
for ln in sys.stdin:
    if ln.startswith('{'):
        j=json.loads(ln); print('  value %.0f frames/s  ms/step %.4f' % (j['value'], j['ms_per_step'])); print('  ', {k: round(v['median'],4) for k,v in j['stages_ms'].items()})
    else: print(ln.rstrip())
"
done
