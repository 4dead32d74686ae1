#!/bin/bash
# voxelizer: scatter finalize + side-stream fill of the voxel tensor, sample groups (BEVVOX_GROUP); camera step: --lanes
mkdir -p gpurun_out
TAG=${1:-r2o}
timeout 900 python -m pytest tests/test_gpu_voxelize.py -m gpu -q -x --timeout 300 2>&1 | tail -4
BEVVOX_GROUP=3 timeout 900 python -m pytest tests/test_gpu_voxelize.py -m gpu -q -x --timeout 300 2>&1 | tail -2
for gsz in 0 8 16; do
  echo "== BEVVOX_GROUP=$gsz"
  BEVVOX_GROUP=$gsz timeout 300 python scripts/lidar_trace.py 32 2>&1 | tail -9
  for n in 8 32; do BEVVOX_GROUP=$gsz timeout 300 python scripts/lidar_probe.py $n 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['sweeps_per_step'],'sweeps', round(d['ms_per_step']*1e3,1),'us graph', round(d['eager_no_sync_ms']*1e3,1),'us eager', 'frac', round(d['frac_of_hbm_peak'],3))"; done
done
for ln in 1 2 4; do
timeout 300 python bench.py --steps 20 --warmup 3 --no-extras --lanes $ln > gpurun_out/bench_${TAG}_l$ln.json 2> gpurun_out/bench_${TAG}_l$ln.err
echo "lanes $ln:"; python scripts/print_stages.py < gpurun_out/bench_${TAG}_l$ln.json 2>&1 | head -1 | cut -c1-60; tail -2 gpurun_out/bench_${TAG}_l$ln.err
done
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:vox_" --launch-skip 14 --launch-count 14 -f -o gpurun_out/prof_${TAG}_lidar python scripts/lidar_probe.py 32 > gpurun_out/prof_${TAG}_lidar.log 2>&1
tail -2 gpurun_out/prof_${TAG}_lidar.log
