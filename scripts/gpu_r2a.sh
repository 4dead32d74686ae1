#!/bin/bash
# round 2, first GPU call: which rig variant matches torch; the new kernels' tests; a short bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv,noheader
timeout 300 python scripts/rig_probe.py 2>&1 | tail -20
timeout 900 python -m pytest tests/test_gpu_nchw_rig.py -m gpu -x -q --timeout 300 2>&1 | tail -25
timeout 300 python bench.py --steps 20 --warmup 3 --no-extras > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err
python scripts/print_stages.py < gpurun_out/bench_r2a.json; tail -5 gpurun_out/bench_r2a.err
