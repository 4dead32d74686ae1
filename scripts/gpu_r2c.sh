#!/bin/bash
mkdir -p gpurun_out
export BEVPOOL_DEBUG=1
timeout 1200 python -m pytest tests/test_gpu_nchw_rig.py -m gpu -q -x --timeout 300 2>&1 | tail -15
timeout 300 python bench.py --steps 20 --warmup 3 --no-extras > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err
python scripts/print_stages.py < gpurun_out/bench_r2c.json; tail -5 gpurun_out/bench_r2c.err
CMD="python bench.py --steps 2 --warmup 1 --no-graph --no-extras"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2c.csv $CMD > gpurun_out/launches_r2c.log 2>&1
python scripts/launch_table.py gpurun_out/launches_r2c.csv 2>&1 | tail -20
KREG="regex:plan_key_rig|fused_backward_col"
timeout 400 ncu --set full --clock-control none --import-source on -k "$KREG" --launch-skip 2 --launch-count 2 -f -o gpurun_out/prof_r2c $CMD > gpurun_out/prof_r2c.log 2>&1
tail -2 gpurun_out/prof_r2c.log
