"""The LiDAR-branch side measurement alone (launch lists / ncu): python scripts/lidar_probe.py [sweeps]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')))
r = bench.run_lidar_side(torch.device('cuda'), float(peaks.get('hbm_gbs', 6650.0)), int(sys.argv[1]) if len(sys.argv) > 1 else 32)
print(json.dumps(r))
