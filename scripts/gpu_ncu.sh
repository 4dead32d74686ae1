#!/bin/bash
# usage: bash scripts/gpu_ncu.sh <tag> ["ENV=.. ENV=.."]  -- launch list + ONE ncu --set full capture of every kernel of a step
mkdir -p gpurun_out
TAG=$1; ENVS=${2:-X=1}
CMD="python bench.py --steps 2 --warmup 1 --no-graph --no-extras"
env $ENVS timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/launches_${TAG}.log 2>&1
KREG='regex:frustum_|cell_sum|plan_key|sort_hist|sort_scatter|bucket_sort|scan_exclusive|pool_forward|fused_backward|grad_rows|transpose_kernel|compact'
env $ENVS timeout 400 ncu --set full --clock-control none --import-source on -k "$KREG" --launch-skip 24 --launch-count 14 -f -o gpurun_out/prof_${TAG}_step $CMD > gpurun_out/prof_${TAG}_step.log 2>&1
tail -2 gpurun_out/prof_${TAG}_step.log
