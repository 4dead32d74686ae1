timeout 900 python -m pytest tests/test_gpu_voxelize.py -m gpu -q -x --timeout 300 2>&1 | tail -3
timeout 300 python scripts/lidar_trace.py 32 2>&1 | tail -10
for n in 8 32; do timeout 300 python scripts/lidar_probe.py $n 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['sweeps_per_step'],'sweeps', round(d['ms_per_step']*1e3,1),'us graph', round(d['eager_no_sync_ms']*1e3,1),'us eager', 'frac', round(d['frac_of_hbm_peak'],3))"; done
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:vox_finalize" --launch-skip 2 --launch-count 1 -f -o gpurun_out/prof_r2y_fin python scripts/lidar_probe.py 32 > gpurun_out/prof_r2y_fin.log 2>&1
