#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r2k}
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 300 2>&1 | tail -8
BEVPOOL_ORDER_CLUSTER=0 timeout 600 python -m pytest tests/test_gpu_run_plan.py -m gpu -q -x --timeout 300 2>&1 | tail -2
BEVPOOL_ORDER_CLUSTER=1 timeout 600 python -m pytest tests/test_gpu_run_plan.py -m gpu -q -x --timeout 300 2>&1 | tail -2
timeout 300 python scripts/lidar_trace.py 32 2>&1 | tail -9
for n in 8 32; do timeout 300 python scripts/lidar_probe.py $n 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['sweeps_per_step'],'sweeps', round(d['ms_per_step']*1e3,1),'us graph', round(d['eager_no_sync_ms']*1e3,1),'us eager', 'frac', round(d['frac_of_hbm_peak'],3))"; done
for cs in 0 2 4 8; do
BEVPOOL_ORDER_CLUSTER=$cs timeout 300 python bench.py --steps 20 --warmup 3 --no-extras > gpurun_out/bench_${TAG}_cs$cs.json 2> gpurun_out/bench_${TAG}_cs$cs.err
echo "cluster $cs:"; python scripts/print_stages.py < gpurun_out/bench_${TAG}_cs$cs.json 2>&1 | head -1 | cut -c1-330; tail -2 gpurun_out/bench_${TAG}_cs$cs.err
done
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:vox_" --launch-skip 12 --launch-count 6 -f -o gpurun_out/prof_${TAG}_lidar python scripts/lidar_probe.py 32 > gpurun_out/prof_${TAG}_lidar.log 2>&1
tail -2 gpurun_out/prof_${TAG}_lidar.log
