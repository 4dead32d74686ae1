#!/bin/bash
# voxelizer: cooperative finalize v2 (A/B by BEVVOX_FIN_COOP), full GPU test suite, bench line
mkdir -p gpurun_out
TAG=${1:-r2q}
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 300 2>&1 | tail -4
BEVVOX_FIN_COOP=0 timeout 900 python -m pytest tests/test_gpu_voxelize.py -m gpu -q -x --timeout 300 2>&1 | tail -2
for c in 0 1; do
  echo "== BEVVOX_FIN_COOP=$c"
  BEVVOX_FIN_COOP=$c timeout 300 python scripts/lidar_trace.py 32 2>&1 | tail -10
  for n in 8 32; do BEVVOX_FIN_COOP=$c timeout 300 python scripts/lidar_probe.py $n 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['sweeps_per_step'],'sweeps', round(d['ms_per_step']*1e3,1),'us graph', round(d['eager_no_sync_ms']*1e3,1),'us eager', 'frac', round(d['frac_of_hbm_peak'],3))"; done
done
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:vox_" --launch-skip 16 --launch-count 8 -f -o gpurun_out/prof_${TAG}_lidar python scripts/lidar_probe.py 32 > gpurun_out/prof_${TAG}_lidar.log 2>&1
tail -2 gpurun_out/prof_${TAG}_lidar.log
