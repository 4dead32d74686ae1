#!/bin/bash
# usage (on a box with 8 GPUs): bash scripts/gpu_pcie.sh  -- raw pinned-copy bandwidth with 1, 2, 4, 8 ranks copying at once
mkdir -p gpurun_out
timeout 120 python scripts/pcie_probe.py > gpurun_out/pcie_n1.json 2>gpurun_out/pcie_n1.err; cat gpurun_out/pcie_n1.json
for n in 2 4 8; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 scripts/pcie_probe.py > gpurun_out/pcie_n$n.json 2>gpurun_out/pcie_n$n.err
  cat gpurun_out/pcie_n$n.json
done
lscpu | grep -i "model name\|numa\|^CPU(s)" | head -6
