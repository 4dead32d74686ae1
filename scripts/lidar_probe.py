"""Times the LiDAR branch alone (for ncu launch lists): python scripts/lidar_probe.py [sweeps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

if __name__ == '__main__':
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    torch.cuda.set_device(0)
    print(bench.run_lidar_side(torch.device('cuda', 0), 6458.4, sweeps=n))
