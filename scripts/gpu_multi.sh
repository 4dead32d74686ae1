#!/bin/bash
# usage: bash scripts/gpu_multi.sh <N> <tag>   -- one box with N GPUs: headline bench + training harness under torchrun
mkdir -p gpurun_out
N=${1:-2}; TAG=${2:-r2m}
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
nvidia-smi topo -m 2>/dev/null | head -14
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $RUN bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_${TAG}_n${N}.json 2> gpurun_out/bench_${TAG}_n${N}.err
python scripts/print_stages.py < gpurun_out/bench_${TAG}_n${N}.json 2>&1 | head -6; tail -3 gpurun_out/bench_${TAG}_n${N}.err
timeout 600 $RUN bench.py --gpus $N --workload train --steps 10 --warmup 3 > gpurun_out/bench_${TAG}_train_n${N}.json 2> gpurun_out/bench_${TAG}_train_n${N}.err
tail -c 1800 gpurun_out/bench_${TAG}_train_n${N}.json; grep -i "nccl\|error" gpurun_out/bench_${TAG}_train_n${N}.err | head -5
timeout 600 $RUN bench.py --gpus $N --workload train --train-cfg cfg2 --steps 10 --warmup 3 > gpurun_out/bench_${TAG}_train_cfg2_n${N}.json 2> gpurun_out/bench_${TAG}_train_cfg2_n${N}.err
tail -c 600 gpurun_out/bench_${TAG}_train_cfg2_n${N}.json
