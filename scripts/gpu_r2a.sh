#!/bin/bash
# round 2, first GPU call: which rig variant matches torch; the new kernels' tests; a short bench; ncu of the new kernels
mkdir -p gpurun_out
export BEVPOOL_DEBUG=1
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv,noheader
timeout 300 python scripts/rig_probe.py 2>&1 | tail -20
timeout 900 python -m pytest tests/test_gpu_nchw_rig.py -m gpu -q --timeout 300 2>&1 | tail -40
timeout 300 python bench.py --steps 20 --warmup 3 --no-extras > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err
python scripts/print_stages.py < gpurun_out/bench_r2a.json; tail -5 gpurun_out/bench_r2a.err
CMD="python bench.py --steps 2 --warmup 1 --no-graph --no-extras"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2a.csv $CMD > gpurun_out/launches_r2a.log 2>&1
KREG='regex:fused_backward_col|grad_rows_tma|frustum_reduce'
timeout 400 ncu --set full --clock-control none --import-source on -k "$KREG" --launch-skip 6 --launch-count 6 -f -o gpurun_out/prof_r2a $CMD > gpurun_out/prof_r2a.log 2>&1
tail -3 gpurun_out/prof_r2a.log
python scripts/launch_table.py gpurun_out/launches_r2a.csv 2>&1 | tail -20
