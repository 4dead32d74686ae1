#!/bin/bash
# like gpu_exp.sh but without the correctness gate (for what-if experiments that break results on purpose)
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
