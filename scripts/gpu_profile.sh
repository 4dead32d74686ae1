#!/bin/bash
# usage: bash scripts/gpu_profile.sh <tag> [kernel-regex ...]
# 1) launch list of one eager bench step (share of each kernel), 2) --set full capture per regex
mkdir -p gpurun_out
TAG=${1:-r1}; shift
CMD="python bench.py --steps 2 --warmup 1 --no-graph --no-extras"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/launches_${TAG}.log 2>&1
for K in "$@"; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 4 -c 2 -f -o gpurun_out/prof_${TAG}_${K} $CMD > gpurun_out/prof_${TAG}_${K}.log 2>&1
done
ls -la gpurun_out
