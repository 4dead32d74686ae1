#!/bin/bash
mkdir -p gpurun_out
export BEVPOOL_DEBUG=1
timeout 900 python -m pytest tests/test_gpu_nchw_rig.py tests/test_gpu_run_plan.py -m gpu -q -x --timeout 300 2>&1 | tail -25
timeout 300 python bench.py --steps 20 --warmup 3 --no-extras > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err
python scripts/print_stages.py < gpurun_out/bench_r2b.json; tail -5 gpurun_out/bench_r2b.err
CMD="python bench.py --steps 2 --warmup 1 --no-graph --no-extras"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2b.csv $CMD > gpurun_out/launches_r2b.log 2>&1
python scripts/launch_table.py gpurun_out/launches_r2b.csv 2>&1 | tail -20
KREG='regex:fused_backward_col'
timeout 400 ncu --set full --clock-control none --import-source on -k "$KREG" --launch-skip 2 --launch-count 1 -f -o gpurun_out/prof_r2b $CMD > gpurun_out/prof_r2b.log 2>&1
tail -2 gpurun_out/prof_r2b.log
