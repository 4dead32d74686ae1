#!/bin/bash
mkdir -p gpurun_out
export BEVPOOL_DEBUG=1
timeout 900 python -m pytest tests/test_gpu_voxelize.py -m gpu -q -x --timeout 300 2>&1 | tail -15
for n in 8 32; do timeout 300 python scripts/lidar_probe.py $n 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['sweeps_per_step'],'sweeps', round(d['ms_per_step']*1e3,1),'us graph', round(d['eager_no_sync_ms']*1e3,1),'us eager', round(d['mmdet3d_style_exact_size_ms']*1e3,1),'us exact', 'frac', round(d['frac_of_hbm_peak'],3))"; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_lidar_r2d.csv python scripts/lidar_probe.py 32 > gpurun_out/launches_lidar_r2d.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_lidar_r2d.csv')) if len(r)>5 and r[0].isdigit()]
# last 8 launches (one graph replay is not visible to ncu per kernel; take the eager tail)
import collections
agg=collections.OrderedDict()
for r in rows[-40:]:
    name=r[4].split('(')[0][:40]; agg.setdefault(name,[]).append(float(r[-1].replace(',','')))
for k,v in agg.items(): print(k, len(v), round(sum(v)/len(v)/1000,1),'us avg')
PY
