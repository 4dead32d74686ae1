#!/bin/bash
# usage: bash scripts/gpu_prof.sh <tag> "<ENV settings>" kernel-regex...   -- ncu --set full of the given kernels (bench step, eager)
mkdir -p gpurun_out
TAG=$1; shift; ENVS=$1; shift
CMD="python bench.py --steps 2 --warmup 1 --no-graph --no-extras"
for K in "$@"; do
  env $ENVS timeout 240 ncu --set full --clock-control none --import-source on -k regex:$K -s 4 -c 1 -f -o gpurun_out/prof_${TAG}_${K} $CMD > gpurun_out/prof_${TAG}_${K}.log 2>&1
  tail -2 gpurun_out/prof_${TAG}_${K}.log
done
ls -la gpurun_out | tail -5
