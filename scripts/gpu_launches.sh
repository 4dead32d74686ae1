#!/bin/bash
# usage: bash scripts/gpu_launches.sh <tag> "<ENV>"  -- per-kernel device time of one eager bench step
mkdir -p gpurun_out
TAG=$1; ENVS=$2
env $ENVS timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 1 --no-graph --no-extras > gpurun_out/launches_${TAG}.log 2>&1
python scripts/launch_table.py gpurun_out/launches_${TAG}.csv
