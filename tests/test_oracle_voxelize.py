"""CPU checks of the voxelizer / VFE / pillar-scatter oracle.  PARITY UNPINNED by the
reference (mmcv/mmdet3d absent): anchored on the hand-derived KAT and on
agreement between the serial C restatement and the vectorised numpy formulation."""
import hashlib
import json
import os

import numpy as np
import pytest

from mm_training_b200 import synthetic
from mm_training_b200.configs import CFG_3
from oracle import voxelize_ref as vz

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')
IMPLS = [vz.hard_voxelize_c, vz.hard_voxelize_numpy]


@pytest.mark.parametrize('impl', IMPLS)
def test_known_answer_appendix_a4(impl):
    kat = json.load(open(os.path.join(GOLDEN, 'voxelize_kat.json')))
    pts = np.asarray(kat['points'], np.float32)
    v, c, n = impl(pts, kat['voxel_size'], kat['point_cloud_range'], kat['max_num_points'],
                   kat['max_voxels'])
    assert c.tolist() == kat['coors'] and n.tolist() == kat['num_points']
    for vid, ids in enumerate(kat['voxel_point_ids']):
        for slot, pid in enumerate(ids):
            exp = pts[pid] if pid >= 0 else np.zeros(3, np.float32)
            assert np.array_equal(v[vid, slot], exp)
    _, _, cb = vz.voxelize_batch_ref([pts], kat['voxel_size'], kat['point_cloud_range'],
                                     kat['max_num_points'], kat['max_voxels'], impl=impl)
    assert cb.tolist() == kat['coors_batch0']


def test_config3_sweep_matches_digest_and_numpy():
    d = json.load(open(os.path.join(GOLDEN, 'voxelize_sweep_digest.json')))
    pts = synthetic.lidar_sweep(CFG_3.points_per_sweep, 5, seed=2)
    assert hashlib.sha256(pts.tobytes()).hexdigest() == d['points_sha256']
    a = vz.hard_voxelize_c(pts, CFG_3.voxel_size, CFG_3.point_cloud_range, CFG_3.max_num_points,
                           CFG_3.max_voxels)
    b = vz.hard_voxelize_numpy(pts, CFG_3.voxel_size, CFG_3.point_cloud_range, CFG_3.max_num_points,
                               CFG_3.max_voxels)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    v, c, n = a
    assert v.shape[0] == d['num_voxels'] == CFG_3.max_voxels          # the cap binds
    assert int(n.sum()) == d['stored_points'] and int(n.max()) == d['max_points']
    assert hashlib.sha256(c.tobytes()).hexdigest() == d['coors_sha256']
    assert hashlib.sha256(n.tobytes()).hexdigest() == d['num_sha256']
    assert hashlib.sha256(v.tobytes()).hexdigest() == d['voxels_sha256']


@pytest.mark.parametrize('seed', range(6))
def test_c_and_numpy_agree_on_random_small(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(0, 400))
    f = int(rng.integers(3, 9))
    pts = rng.uniform(-1.5, 5.5, size=(n, f)).astype(np.float32)
    if n > 10:                                   # duplicates and on-boundary points
        pts[5] = pts[2]
        pts[7, :3] = [0.0, 0.0, 0.0]
        pts[8, :3] = [4.0, 1.0, 0.5]
    args = ([0.5, 0.25, 1.0], [0, 0, 0, 4, 4, 2], int(rng.integers(1, 5)), int(rng.integers(1, 40)))
    a, b = vz.hard_voxelize_c(pts, *args), vz.hard_voxelize_numpy(pts, *args)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


@pytest.mark.parametrize('impl', IMPLS)
def test_edge_cases(impl):
    rngargs = ([1, 1, 1], [0, 0, 0, 4, 4, 1])
    # empty input
    v, c, n = impl(np.zeros((0, 4), np.float32), *rngargs, 3, 5)
    assert v.shape == (0, 3, 4) and c.shape == (0, 3) and n.shape == (0,)
    # all points in one voxel: first max_points kept in order
    pts = np.tile(np.array([[0.5, 0.5, 0.5, 0]], np.float32), (50, 1))
    pts[:, 3] = np.arange(50)
    v, c, n = impl(pts, *rngargs, 4, 5)
    assert n.tolist() == [4] and v[0, :, 3].tolist() == [0, 1, 2, 3]
    # exactly max_voxels + 1 distinct voxels: last one dropped
    pts = np.array([[x + 0.5, y + 0.5, 0.5] for y in range(2) for x in range(3)], np.float32)
    v, c, n = impl(pts, *rngargs, 2, 5)
    assert v.shape[0] == 5 and c[-1].tolist() == [0, 1, 1]
    # bounds: min is inside, max is outside (floor semantics); negative epsilon is outside
    pts = np.array([[0, 0, 0], [4, 0, 0], [-1e-6, 0, 0], [3.9999, 3.9999, 0.9999]], np.float32)
    v, c, n = impl(pts, *rngargs, 2, 5)
    assert c.tolist() == [[0, 0, 0], [0, 3, 3]]


def test_vfe_and_pillar_scatter_oracle():
    rng = np.random.default_rng(0)
    pts = synthetic.lidar_sweep(5000, 8, seed=5)
    vox, num, coors = vz.voxelize_batch_ref([pts, pts[::-1].copy()], CFG_3.voxel_size,
                                            CFG_3.point_cloud_range, 15, 25000)
    feats = vz.hard_simple_vfe_ref(vox, num, 5)
    assert feats.shape == (vox.shape[0], 5)
    m = 17
    manual = vox[m, :num[m], :5].sum(0) / num[m]
    assert np.allclose(feats[m], manual, rtol=1e-6)
    canvas = vz.pillar_scatter_ref(feats, coors, 2, (1, 256, 2048))
    assert canvas.shape == (2, 5, 256, 2048)
    b, z, y, x = coors[m]
    assert np.array_equal(canvas[b, :, y, x], feats[m])
    assert int((np.abs(canvas).sum(1) > 0).sum()) <= vox.shape[0]
    g = rng.random(canvas.shape).astype(np.float32)
    gb = vz.pillar_scatter_backward_ref(g, coors, (1, 256, 2048))
    assert np.array_equal(gb[m], g[b, :, y, x])


def test_reciprocal_cell_index_rule_equals_the_exact_division():
    """csrc/voxelize.cu::point_to_cell_fast in numpy float32: q' = (p - min) * fl(1 / v) decides floor(fl((p - min) / v))
    whenever q' is farther than 1e-3 from every integer and |q'| < 4096; otherwise the exact expression is used.  Checked
    on random coordinates, on coordinates placed within a few ulps of every cell boundary, and for awkward voxel sizes."""
    rng = np.random.default_rng(0)
    for v in (0.2, 0.1, 0.05, 0.075, 0.8, 1.0 / 3.0, 0.16, 2.5):
        v32 = np.float32(v)
        inv = np.float32(1.0) / v32
        lo = np.float32(-204.8 if v < 1 else -37.3)
        cells = np.arange(0, 4096, dtype=np.float64)
        edges = (np.float64(lo) + cells * np.float64(v32)).astype(np.float32)
        near = np.concatenate([np.nextafter(edges, np.float32(np.inf)), np.nextafter(edges, np.float32(-np.inf)), edges])
        for k in range(1, 6):                                        # a few more ulps around every boundary
            up = near.copy()
            for _ in range(k):
                up = np.nextafter(up, np.float32(np.inf))
            near = np.concatenate([near, up])
        p = np.concatenate([near, rng.uniform(float(lo) - 5, float(lo) + 4096 * v + 5, 2_000_000).astype(np.float32)])
        t = (p - lo).astype(np.float32)
        exact = np.floor((t / v32).astype(np.float32))
        q = (t * inv).astype(np.float32)
        fast_ok = (np.abs(q - np.rint(q)) > np.float32(1e-3)) & (np.abs(q) < np.float32(4096.0))
        assert np.array_equal(np.floor(q)[fast_ok], exact[fast_ok]), v
        ordinary = np.arange(p.size) >= near.size                    # the uniformly drawn coordinates
        ordinary &= (q > 0) & (q < 4000)
        assert fast_ok[ordinary].mean() > 0.99                       # (the exact path is the rare one on ordinary data)
