#!/bin/bash
# full GPU suite on the current build + stage A depth-split A/B
mkdir -p gpurun_out
TAG=${1:-r2u}
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 300 2>&1 | tail -4
for v in 0 1; do
BEVPOOL_RUN_DSPLIT=$v timeout 300 python bench.py --steps 20 --warmup 3 --no-extras > gpurun_out/bench_${TAG}_d$v.json 2> gpurun_out/bench_${TAG}_d$v.err
echo "dsplit $v:"; python scripts/print_stages.py < gpurun_out/bench_${TAG}_d$v.json 2>&1 | head -1 | cut -c1-330; tail -2 gpurun_out/bench_${TAG}_d$v.err
done
