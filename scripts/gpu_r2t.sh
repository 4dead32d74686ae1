#!/bin/bash
# voxelizer: finalize in parts with the canvas pass on a side stream (A/B by BEVVOX_FIN_PARTS)
mkdir -p gpurun_out
TAG=${1:-r2t}
for p in 3 1; do BEVVOX_FIN_PARTS=$p timeout 900 python -m pytest tests/test_gpu_voxelize.py -m gpu -q -x --timeout 300 2>&1 | tail -2; done
for p in 1 2 3 4; do
  echo "== BEVVOX_FIN_PARTS=$p"
  for n in 8 32; do BEVVOX_FIN_PARTS=$p timeout 300 python scripts/lidar_probe.py $n 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['sweeps_per_step'],'sweeps', round(d['ms_per_step']*1e3,1),'us graph', round(d['eager_no_sync_ms']*1e3,1),'us eager', 'frac', round(d['frac_of_hbm_peak'],3))"; done
done
