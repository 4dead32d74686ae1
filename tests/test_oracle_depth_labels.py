"""CPU tests of the depth-label oracle (oracle/depth_labels_ref.py): the restatement of the reference's torch ops
(exps/mm_training_aim.py:115-215) against a hand-derived known-answer case, against the committed golden fixture,
and against the float32 emulation of the CUDA kernel's arithmetic order."""
import os

import numpy as np
import torch

from oracle import depth_labels_ref as dl

HERE = os.path.dirname(os.path.abspath(__file__))
DB = (2.0, 58.0, 0.5)
D = 112


def test_known_answer_last_point_wins_and_min_pool():
    # one camera looking along +x of the ego frame (identity BDA), f = 16, 32 x 32 image, 16 x 16 blocks -> 2 x 2 cells
    r0 = torch.tensor([[0, 0, 1.], [-1, 0, 0], [0, -1, 0]])
    cam2ego = torch.eye(4)
    cam2ego[:3, :3] = r0
    ext = torch.linalg.inv(cam2ego)[None, None, None]
    k = torch.eye(4)
    k[0, 0] = k[1, 1] = 16.0
    k[0, 2] = k[1, 2] = 16.0
    intr = k[None, None, None]
    bda = torch.eye(4)[None]
    # ego (x fwd, y left, z up) -> camera (x right = -y, y down = -z, z fwd = x)
    pts = torch.tensor([[10.0, 0.0, 0.0],      # centre pixel (16, 16): block (1, 1), depth 10
                        [20.0, 0.0, 0.0],      # same pixel, later in the cloud -> overwrites 10 with 20
                        [8.0, 2.0, 2.0],       # u = 16 - 4 = 12, v = 12: block (0, 0), depth 8
                        [4.0, 1.0, 1.0],       # same pixel (12, 12), later -> depth 4 wins
                        [30.0, -7.5, 1.875],   # u = 20, v = 15: block (0, 1), depth 30
                        [0.5, 0.0, 0.0],       # depth <= 1: dropped
                        [10.0, 20.0, 0.0]])    # u < 1: dropped
    pts = torch.cat([pts, torch.zeros(len(pts), 2)], 1)        # (Np, 5) like the loader's [x, y, z, intensity, t]
    onehot, bins = dl.depth_labels_torch([pts], ext, intr, bda, (32, 32), 16, DB, D)
    # bin = trunc((d - 1.5) / 0.5): 4 -> 5, 30 -> 57, 20 -> 37; block (1, 0) is empty -> 1e5 -> out of range -> 0
    assert bins.tolist() == [5, 57, 0, 37]
    assert onehot.shape == (4, D) and onehot.sum() == 4 and onehot[0, 5] == 1 and onehot[2, 0] == 1
    _, bins_e = dl.depth_labels_exact([pts], ext, intr, bda, (32, 32), 16, DB, D)
    assert torch.equal(bins_e, bins)


def test_exact_emulation_equals_torch_ops_when_products_are_exact():
    case = dl.synthetic_case(batch=2, sweeps=2, cams=2, num_points=30000, image_hw=(64, 128), seed=3, exact_products=True)
    a, ab = dl.depth_labels_torch(*case, (64, 128), 16, DB, D)
    b, bb = dl.depth_labels_exact(*case, (64, 128), 16, DB, D)
    assert torch.equal(ab, bb) and torch.equal(a, b)
    assert int((ab > 0).sum()) > ab.numel() // 4          # the case actually hits most cells


def test_exact_emulation_vs_torch_ops_on_general_rigs():
    case = dl.synthetic_case(batch=2, sweeps=1, cams=2, num_points=60000, image_hw=(128, 256), seed=5)
    _, ab = dl.depth_labels_torch(*case, (128, 256), 16, DB, D)
    _, bb = dl.depth_labels_exact(*case, (128, 256), 16, DB, D)
    # accumulation order of torch's 4x4 @ 4xN matmul vs the kernel's left-to-right order: a last-bit difference can
    # move a point across a pixel / mask / bin boundary.  Bound stated here: at most 1 cell in 1000.
    assert float((ab != bb).float().mean()) <= 1e-3
    assert int((ab > 0).sum()) > ab.numel() // 3


def test_golden_fixture():
    z = np.load(os.path.join(HERE, 'golden', 'depth_labels_golden.npz'))
    case = dl.synthetic_case(batch=2, sweeps=1, cams=2, num_points=20000, image_hw=(64, 128), seed=int(z['seed']), exact_products=True)
    _, bins = dl.depth_labels_torch(*case, (64, 128), 16, DB, D)
    assert np.array_equal(bins.numpy(), z['bins'])
    _, bins_e = dl.depth_labels_exact(*case, (64, 128), 16, DB, D)
    assert np.array_equal(bins_e.numpy(), z['bins'])
