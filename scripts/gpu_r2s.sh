#!/bin/bash
# backward: static-bin loop (A/B by BEVPOOL_BW_STATIC)
mkdir -p gpurun_out
TAG=${1:-r2s}
timeout 1200 python -m pytest tests/test_gpu_nchw_rig.py tests/test_gpu_voxel_pool.py tests/test_gpu_run_plan.py tests/test_gpu_host_pipeline.py -m gpu -q -x --timeout 300 2>&1 | tail -4
for v in 0 1; do
BEVPOOL_BW_STATIC=$v timeout 300 python bench.py --steps 20 --warmup 3 --no-extras > gpurun_out/bench_${TAG}_s$v.json 2> gpurun_out/bench_${TAG}_s$v.err
echo "static $v:"; python scripts/print_stages.py < gpurun_out/bench_${TAG}_s$v.json 2>&1 | head -1 | cut -c1-330; tail -2 gpurun_out/bench_${TAG}_s$v.err
done
CMD="python bench.py --steps 2 --warmup 1 --no-graph --no-extras"
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:fused_backward" --launch-skip 2 --launch-count 2 -f -o gpurun_out/prof_${TAG}_bwd $CMD > gpurun_out/prof_${TAG}_bwd.log 2>&1
tail -2 gpurun_out/prof_${TAG}_bwd.log
