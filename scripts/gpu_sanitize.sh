#!/bin/bash
# compute-sanitizer over the GPU parity tests: memcheck on everything, racecheck on the fused kernels
mkdir -p gpurun_out
timeout 700 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/memcheck_all.log 2>&1; echo memcheck rc=$?
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_gpu_run_plan.py tests/test_gpu_voxel_pool.py -x -q -k "run_plan_and_forward or camera_rig or fused_forward_backward_random" --timeout 450 > gpurun_out/racecheck_fused.log 2>&1; echo racecheck rc=$?
tail -4 gpurun_out/memcheck_all.log gpurun_out/racecheck_fused.log
