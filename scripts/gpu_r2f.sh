#!/bin/bash
# round evidence at HEAD + LiDAR trace
bash scripts/gpu_round.sh r2f
timeout 300 python scripts/lidar_trace.py 32 2>&1 | tail -14
timeout 300 python scripts/lidar_trace.py 8 2>&1 | tail -3
for n in 8 32; do timeout 300 python scripts/lidar_probe.py $n 2>&1 | tail -1 > gpurun_out/lidar_r2f_$n.json; python -c "
import json,sys
d=json.load(open('gpurun_out/lidar_r2f_$n.json'))
print(d['sweeps_per_step'],'sweeps', round(d['ms_per_step']*1e3,1),'us graph', round(d['eager_no_sync_ms']*1e3,1),'us eager', round(d['mmdet3d_style_exact_size_ms']*1e3,1),'us exact', 'frac', round(d['frac_of_hbm_peak'],3))"; done
