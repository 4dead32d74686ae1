"""Generates tests/golden/depth_labels_golden.npz with the torch restatement of the reference's depth-label functions
(oracle/depth_labels_ref.py::depth_labels_torch, exps/mm_training_aim.py:115-215).  Run from the repo root:
    python tests/golden/make_depth_labels_golden.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import depth_labels_ref as dl  # noqa: E402

seed = 11
case = dl.synthetic_case(batch=2, sweeps=1, cams=2, num_points=20000, image_hw=(64, 128), seed=seed, exact_products=True)
_, bins = dl.depth_labels_torch(*case, (64, 128), 16, (2.0, 58.0, 0.5), 112)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'depth_labels_golden.npz'), seed=seed,
                    bins=bins.numpy())
print('cells', bins.numel(), 'labelled', int((bins > 0).sum()))
