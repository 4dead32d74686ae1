#!/bin/bash
# usage: bash scripts/gpu_round.sh <tag> [kernel-regex ...]   -- smoke, tests, bench, launch list, ncu --set full per regex
# every step runs under its own timeout: a hung kernel must not eat the GPU budget
mkdir -p gpurun_out
TAG=${1:-r1}; shift
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python -m pytest tests -m gpu -x -q --timeout 120 2>&1 | tail -15
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; tail -c 3500 gpurun_out/bench_${TAG}.json; tail -5 gpurun_out/bench_${TAG}.err
CMD="python bench.py --steps 2 --warmup 1 --no-graph --no-extras"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/launches_${TAG}.log 2>&1
for K in "$@"; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$K -s 4 -c 2 -f -o gpurun_out/prof_${TAG}_${K} $CMD > gpurun_out/prof_${TAG}_${K}.log 2>&1
done
ls -la gpurun_out | tail -8
