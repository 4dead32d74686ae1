#!/bin/bash
# usage: bash scripts/gpu_evidence.sh <tag> [tests]  -- the evidence set of a round in one call: smoke, (GPU tests), bench with extras
# (both arms), launch lists with DRAM bytes (camera step + LiDAR branch), full ncu captures of both kernel sets
mkdir -p gpurun_out
TAG=${1:-r2r}
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv,noheader
timeout 240 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
if [ "$2" == "tests" ]; then timeout 1200 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -6; fi
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
python scripts/print_stages.py < gpurun_out/bench_${TAG}.json 2>&1 | tail -40; tail -5 gpurun_out/bench_${TAG}.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${TAG}_reference.json 2>> gpurun_out/bench_${TAG}.err
timeout 300 python scripts/lidar_trace.py 32 2>&1 | tail -10
CMD="python bench.py --steps 2 --warmup 1 --no-graph --no-extras"
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/launches_${TAG}.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_lidar_${TAG}.csv python scripts/lidar_probe.py 32 > gpurun_out/launches_lidar_${TAG}.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:vox_|scatter_" --launch-skip 16 --launch-count 8 -f -o gpurun_out/prof_${TAG}_lidar python scripts/lidar_probe.py 32 > gpurun_out/prof_${TAG}_lidar.log 2>&1
tail -2 gpurun_out/prof_${TAG}_lidar.log
KREG='regex:frustum_|plan_key|run_csr|run_place|run_finish|pool_forward|fused_backward|grad_rows'
timeout 400 ncu --set full --clock-control none --import-source on -k "$KREG" --launch-skip 20 --launch-count 10 -f -o gpurun_out/prof_${TAG}_step $CMD > gpurun_out/prof_${TAG}_step.log 2>&1
tail -2 gpurun_out/prof_${TAG}_step.log
