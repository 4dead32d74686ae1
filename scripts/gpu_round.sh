#!/bin/bash
# usage: bash scripts/gpu_round.sh <tag> [skip-tests]
# smoke, GPU tests, bench (both arms), launch list of one eager step, ONE ncu --set full capture holding
# every kernel of the step.  Every stage runs under its own timeout: a hung kernel must not eat the budget.
mkdir -p gpurun_out
TAG=${1:-r1}
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv,noheader
timeout 240 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
if [ "$2" != "skip-tests" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q --timeout 180 2>&1 | tail -15
fi
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
tail -c 4000 gpurun_out/bench_${TAG}.json; tail -5 gpurun_out/bench_${TAG}.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${TAG}_reference.json 2>> gpurun_out/bench_${TAG}.err
tail -c 1500 gpurun_out/bench_${TAG}_reference.json
CMD="python bench.py --steps 2 --warmup 1 --no-graph --no-extras"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/launches_${TAG}.log 2>&1
KREG='regex:frustum_|cell_sum|plan_key|sort_hist|sort_scatter|bucket_sort|scan_exclusive|pool_forward|fused_backward|grad_rows|transpose_kernel|column_'
timeout 400 ncu --set full --clock-control none --import-source on -k "$KREG" --launch-skip 24 --launch-count 16 -f -o gpurun_out/prof_${TAG}_step $CMD > gpurun_out/prof_${TAG}_step.log 2>&1
tail -3 gpurun_out/prof_${TAG}_step.log
ls -la gpurun_out | tail -8
