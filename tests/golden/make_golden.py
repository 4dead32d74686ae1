"""Regenerates the committed golden fixtures.  Run from the repo root:

    python tests/golden/make_golden.py

* ``voxel_pool_reftest.npz`` -- the reference's only known-answer test
  (``test/test_ops/test_voxel_pooling.py:15-37``): the inputs are re-created from its
  seeds, the golden is its own python double loop (``oracle.voxel_pool_ref.python_loop_golden``
  restates it verbatim).  The full (2,80,128,128) golden is 10 MB, so the fixture keeps
  the kept mask digest, per-channel sums, the occupied-cell count and 256 probe cells.
* ``voxelize_kat.json`` -- the hand-derived known-answer test of SURVEY.md Appendix A.4
  (not produced by code: typed in from the derivation).
* ``voxelize_sweep_digest.json`` -- digest of our C restatement on the config-3 synthetic
  sweep (NOT a reference output: mmcv is not available; it only guards against drift).
"""
import hashlib
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import voxel_pool_ref as vp          # noqa: E402
from oracle import voxelize_ref as vz            # noqa: E402
from mm_training_b200 import synthetic           # noqa: E402
from mm_training_b200.configs import CFG_3       # noqa: E402


def main():
    geom, feats = vp.reference_test_inputs()
    gold = vp.python_loop_golden(geom, feats, (128, 128, 1)).contiguous()   # (2,80,128,128)
    kept, lin, _ = vp.cell_index_ref(geom.int(), (128, 128, 1))
    rng = np.random.default_rng(0)
    occ = torch.unique(lin[kept])
    probes = occ[torch.from_numpy(rng.choice(occ.numel(), 256, replace=False))]
    rows = gold.permute(0, 2, 3, 1).reshape(-1, 80)[probes]
    np.savez_compressed(
        os.path.join(HERE, 'voxel_pool_reftest.npz'),
        kept_sha256=np.frombuffer(hashlib.sha256(kept.numpy().tobytes()).digest(), dtype=np.uint8),
        kept_count=np.int64(kept.sum().item()),
        occupied_cells=np.int64(occ.numel()),
        channel_sums=gold.double().sum(dim=(0, 2, 3)).numpy(),
        probe_cells=probes.numpy(), probe_rows=rows.numpy())

    pts = synthetic.lidar_sweep(CFG_3.points_per_sweep, 5, seed=2)
    v, c, n = vz.hard_voxelize_c(pts, CFG_3.voxel_size, CFG_3.point_cloud_range,
                                 CFG_3.max_num_points, CFG_3.max_voxels)
    digest = dict(num_voxels=int(v.shape[0]), stored_points=int(n.sum()), max_points=int(n.max()),
                  coors_sha256=hashlib.sha256(c.tobytes()).hexdigest(),
                  num_sha256=hashlib.sha256(n.tobytes()).hexdigest(),
                  voxels_sha256=hashlib.sha256(v.tobytes()).hexdigest(),
                  points_sha256=hashlib.sha256(pts.tobytes()).hexdigest())
    with open(os.path.join(HERE, 'voxelize_sweep_digest.json'), 'w') as f:
        json.dump(digest, f, indent=1)
    print(digest)


if __name__ == '__main__':
    main()
