#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_voxelize.py -m gpu -q -x --timeout 300 2>&1 | tail -3
timeout 300 python scripts/lidar_trace.py 32 2>&1 | tail -8
timeout 300 python scripts/lidar_trace.py 8 2>&1 | tail -2
