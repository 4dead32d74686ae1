"""Kernel + memset durations of one eager LiDAR step from the CUPTI trace (torch.profiler): unlike ncu, nothing is
serialised or cache-flushed.  python scripts/lidar_trace.py [sweeps]"""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mm_training_b200 import synthetic  # noqa: E402
from mm_training_b200.configs import CFG_3  # noqa: E402
from mm_training_b200.ops.voxelize import Voxelization, voxelize  # noqa: E402

sweeps = int(sys.argv[1]) if len(sys.argv) > 1 else 32
v = CFG_3
clouds = [torch.from_numpy(synthetic.lidar_sweep(v.points_per_sweep, v.num_point_features, seed=2 + i)).cuda() for i in range(sweeps)]
layer = Voxelization(list(v.voxel_size), list(v.point_cloud_range), v.max_num_points, v.max_voxels).eval()
run = lambda: voxelize(clouds, layer, mean_features=v.vfe_features, padded=True, scatter=True)
for _ in range(3):
    run()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5):
        run()
    torch.cuda.synchronize()
rows = {}
for e in prof.events():
    if e.device_type.name == 'CUDA':
        rows.setdefault(e.name[:60], []).append(e.device_time)
tot = 0
for k, t in rows.items():
    print(f'{k:62s} n={len(t):3d}  avg {sum(t)/len(t):8.1f} us')
    tot += sum(t) / 5
print('sum per step', round(tot, 1), 'us')
