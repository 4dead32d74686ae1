#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r2h}
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -15
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
python scripts/print_stages.py < gpurun_out/bench_${TAG}.json 2>&1 | tail -30; tail -5 gpurun_out/bench_${TAG}.err
BEVPOOL_COLD_OVERLAP=0 timeout 300 python bench.py --steps 20 --warmup 3 --no-extras > gpurun_out/bench_${TAG}_nooverlap.json 2>> gpurun_out/bench_${TAG}.err
python scripts/print_stages.py < gpurun_out/bench_${TAG}_nooverlap.json 2>&1 | head -2
timeout 300 python scripts/lidar_trace.py 32 2>&1 | tail -9
timeout 300 python scripts/lidar_trace.py 8 2>&1 | tail -8
BEVVOX_CANVAS_OVERLAP=0 timeout 300 python scripts/lidar_probe.py 32 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('no canvas overlap:', round(d['ms_per_step']*1e3,1),'us graph', 'frac', round(d['frac_of_hbm_peak'],3))"
timeout 600 python bench.py --workload train --train-cfg cfg2 --steps 10 --warmup 3 > gpurun_out/bench_${TAG}_train_cfg2.json 2> gpurun_out/bench_${TAG}_train_cfg2.err; tail -c 2500 gpurun_out/bench_${TAG}_train_cfg2.json; tail -3 gpurun_out/bench_${TAG}_train_cfg2.err
timeout 600 python bench.py --workload train --steps 10 --warmup 3 > gpurun_out/bench_${TAG}_train.json 2> gpurun_out/bench_${TAG}_train.err; tail -c 2500 gpurun_out/bench_${TAG}_train.json; tail -3 gpurun_out/bench_${TAG}_train.err
CMD="python bench.py --steps 2 --warmup 1 --no-graph --no-extras"
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/launches_${TAG}.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_lidar_${TAG}.csv python scripts/lidar_probe.py 32 > gpurun_out/launches_lidar_${TAG}.log 2>&1
