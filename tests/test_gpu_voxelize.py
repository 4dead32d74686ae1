"""GPU parity tests of the LiDAR branch: hard voxelizer, fused VFE mean, dynamic voxelization and
pillar scatter against the CPU oracle (serial C restatement of mmcv 1.7.0 + numpy mirror).
Integer outputs (coors, num_points) and the copied point rows must match bit-exactly.
PARITY UNPINNED by the reference (mmcv / mmdet3d absent): see oracle/voxelize_ref.py."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from mm_training_b200 import synthetic
from mm_training_b200.configs import CFG_3
from mm_training_b200.ops.voxelize import (HardSimpleVFE, PointPillarsScatter, Voxelization, dynamic_voxelize,
                                           hard_voxelize_batch, pillar_scatter, voxelize)
from oracle import voxelize_ref as vz

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')
DEV = 'cuda'


def _check_batch(points_list, voxel_size, rng, T, M, mean_features=0):
    gpu = hard_voxelize_batch([torch.from_numpy(p).to(DEV) for p in points_list], voxel_size, rng, T, M,
                              mean_features)
    voxels, num, coors, base, mean = gpu
    rv, rn, rc = vz.voxelize_batch_ref(points_list, voxel_size, rng, T, M)
    assert np.array_equal(coors.cpu().numpy(), rc)
    assert np.array_equal(num.cpu().numpy(), rn)
    assert np.array_equal(voxels.cpu().numpy(), rv)
    counts = [int((rc[:, 0] == b).sum()) for b in range(len(points_list))]
    assert base.tolist() == np.concatenate([[0], np.cumsum(counts)]).tolist()
    if mean_features:
        ref_mean = vz.hard_simple_vfe_ref(rv, rn, mean_features)
        assert np.allclose(mean.cpu().numpy(), ref_mean, rtol=1e-6, atol=1e-7)
    return gpu


def test_known_answer_appendix_a4():
    kat = json.load(open(os.path.join(GOLDEN, 'voxelize_kat.json')))
    pts = torch.tensor(kat['points'], dtype=torch.float32, device=DEV)
    layer = Voxelization(kat['voxel_size'], kat['point_cloud_range'], kat['max_num_points'], kat['max_voxels']).eval()
    voxels, coors, num = layer(pts)
    assert coors.tolist() == kat['coors'] and num.tolist() == kat['num_points']
    assert coors.dtype == torch.int32 and num.dtype == torch.int32
    for vid, ids in enumerate(kat['voxel_point_ids']):
        for slot, pid in enumerate(ids):
            exp = pts[pid] if pid >= 0 else torch.zeros(3, device=DEV)
            assert torch.equal(voxels[vid, slot], exp)
    v2, n2, cb = voxelize([pts], layer)                       # mmdet3d order: voxels, num_points, coors
    assert cb.tolist() == kat['coors_batch0'] and torch.equal(v2, voxels) and torch.equal(n2, num)


@pytest.mark.parametrize('F', [5, 8])
def test_config3_sweep_bit_exact(F):
    pts = synthetic.lidar_sweep(CFG_3.points_per_sweep, F, seed=2)
    voxels, num, coors, base, mean = _check_batch([pts], CFG_3.voxel_size, CFG_3.point_cloud_range,
                                                  CFG_3.max_num_points, CFG_3.max_voxels, mean_features=5)
    assert voxels.shape == (CFG_3.max_voxels, CFG_3.max_num_points, F)       # the cap binds
    if F == 5:
        d = json.load(open(os.path.join(GOLDEN, 'voxelize_sweep_digest.json')))
        assert hashlib.sha256(coors[:, 1:].contiguous().cpu().numpy().tobytes()).hexdigest() == d['coors_sha256']
        assert hashlib.sha256(num.cpu().numpy().tobytes()).hexdigest() == d['num_sha256']
        assert hashlib.sha256(voxels.cpu().numpy().tobytes()).hexdigest() == d['voxels_sha256']
    # run-to-run bit stability
    again = hard_voxelize_batch([torch.from_numpy(pts).to(DEV)], CFG_3.voxel_size, CFG_3.point_cloud_range,
                                CFG_3.max_num_points, CFG_3.max_voxels)
    assert torch.equal(again[0], voxels) and torch.equal(again[1], num) and torch.equal(again[2], coors)


def test_batched_ragged_samples():
    clouds = [synthetic.lidar_sweep(n, 5, seed=10 + i) for i, n in enumerate([30000, 1, 70001, 12345])]
    clouds.insert(2, np.zeros((0, 5), np.float32))            # an empty sample in the middle
    _check_batch(clouds, CFG_3.voxel_size, CFG_3.point_cloud_range, 15, 25000, mean_features=5)
    _check_batch(clouds, CFG_3.voxel_size, CFG_3.point_cloud_range, 3, 500)     # both caps bind


@pytest.mark.parametrize('seed', range(8))
def test_random_small_grids(seed):
    rng = np.random.default_rng(seed)
    clouds = []
    for _ in range(int(rng.integers(1, 4))):
        n = int(rng.integers(0, 3000))
        p = rng.uniform(-1.5, 5.5, size=(n, int(4))).astype(np.float32)
        if n > 20:
            p[5] = p[2]                                       # duplicate point
            p[7, :3] = [0.0, 0.0, 0.0]                        # on the lower bound: inside
            p[8, :3] = [4.0, 1.0, 0.5]                        # on the upper bound: outside
            p[9, :3] = [-1e-6, 1.0, 0.5]
        clouds.append(p)
    _check_batch(clouds, [0.5, 0.25, 1.0], [0, 0, 0, 4, 4, 2], int(rng.integers(1, 6)), int(rng.integers(1, 300)))


def test_adversarial_sets():
    geom = ([1, 1, 1], [0, 0, 0, 4, 4, 1])
    one = np.tile(np.array([[0.5, 0.5, 0.5, 0]], np.float32), (50000, 1))        # all points in one voxel
    one[:, 3] = np.arange(50000)
    v, n, c, _, _ = _check_batch([one], *geom, 15, 5)
    assert n.tolist() == [15] and v[0, :, 3].tolist() == list(range(15))
    cap = np.array([[x + 0.5, y + 0.5, 0.5] for y in range(2) for x in range(3)], np.float32)
    v, n, c, _, _ = _check_batch([cap], *geom, 2, 5)                             # exactly max_voxels + 1 cells
    assert v.shape[0] == 5
    outside = np.full((1000, 3), 9.0, np.float32)                                  # nothing in range
    v, n, c, base, _ = _check_batch([outside, outside], *geom, 2, 5)
    assert v.shape[0] == 0 and base.tolist() == [0, 0, 0]
    nan = np.array([[np.nan, 0.5, 0.5], [0.5, 0.5, 0.5], [np.inf, 0.5, 0.5], [1e30, 0.5, 0.5]], np.float32)
    v, n, c, _, _ = hard_voxelize_batch([torch.from_numpy(nan).to(DEV)], *geom, 2, 5)
    assert c.tolist() == [[0, 0, 0, 0]] and n.tolist() == [1]                      # NaN / inf dropped (documented)


@pytest.mark.parametrize('T', [1, 3, 4, 8, 15, 16, 20, 35])
def test_slot_lists_under_contention(T):
    # every row length of the claim kernel (1..4 quads in registers, longer rows in a loop): 120 k points shuffled over
    # 60 cells (2000 points per voxel arriving in random order) plus a sparse background
    rng = np.random.default_rng(100 + T)
    hot = np.stack([rng.integers(0, 10, 120000) + rng.random(120000) * 0.999,
                    rng.integers(0, 6, 120000) + rng.random(120000) * 0.999, np.full(120000, 0.5)], 1)
    cold = np.stack([rng.uniform(0, 64, 30000), rng.uniform(0, 64, 30000), np.full(30000, 0.5)], 1)
    pts = np.concatenate([hot, cold]).astype(np.float32)
    pts = np.concatenate([pts[rng.permutation(len(pts))], np.arange(len(pts), dtype=np.float32)[:, None]], 1)
    _check_batch([pts, pts[::-1].copy()], [1, 1, 1], [0, 0, 0, 64, 64, 1], T, 3000, mean_features=4)


@pytest.mark.parametrize('hash_path', [False, True])
def test_dense_and_hash_tables_agree(hash_path, monkeypatch):
    # grids up to 2^26 cells use a dense first-point table, larger ones a hash table: same outputs
    monkeypatch.setenv('BEVVOX_FORCE_HASH', '1' if hash_path else '0')
    clouds = [synthetic.lidar_sweep(n, 5, seed=20 + i) for i, n in enumerate([50000, 20000])]
    _check_batch(clouds, CFG_3.voxel_size, CFG_3.point_cloud_range, 15, 25000, mean_features=5)
    _check_batch(clouds, CFG_3.voxel_size, CFG_3.point_cloud_range, 2, 300, mean_features=3)


def test_padded_no_sync_and_fused_scatter():
    # padded=True: no voxel-count read-back (CUDA-graph capturable); scatter=True: voxelize -> VFE mean -> dense canvas
    clouds_np = [synthetic.lidar_sweep(n, 5, seed=30 + i) for i, n in enumerate([60000, 1, 45000])]
    clouds = [torch.from_numpy(p).to(DEV) for p in clouds_np]
    T, M = 15, 25000
    exact = hard_voxelize_batch(clouds, CFG_3.voxel_size, CFG_3.point_cloud_range, T, M, mean_features=5)
    voxels, num, coors, base, mean, canvas = hard_voxelize_batch(clouds, CFG_3.voxel_size, CFG_3.point_cloud_range, T, M,
                                                                 mean_features=5, padded=True, scatter=True)
    assert base.is_cuda and voxels.shape[0] == 3 * M
    n = int(base[-1])
    assert n == exact[0].shape[0]
    assert torch.equal(voxels[:n], exact[0]) and torch.equal(num[:n], exact[1]) and torch.equal(coors[:n], exact[2])
    assert torch.equal(mean[:n], exact[4])
    assert float(voxels[n:].abs().sum()) == 0.0 and int(num[n:].abs().sum()) == 0      # padding rows are zero
    gx, gy, gz = 2048, 256, 1
    ref = vz.pillar_scatter_ref(exact[4].cpu().numpy(), exact[2].cpu().numpy(), 3, (gz, gy, gx))
    assert canvas.shape == (3, 5 * gz, gy, gx) and np.array_equal(canvas.cpu().numpy(), ref)
    # the separate scatter call, both implementations
    a = pillar_scatter(exact[4], exact[2], 3, (gz, gy, gx), unique_coors=True)
    b = pillar_scatter(exact[4], exact[2], 3, (gz, gy, gx))
    assert torch.equal(a, canvas) and torch.equal(b, canvas)
    # capturable: the whole padded call inside a CUDA graph gives the same bits
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        hard_voxelize_batch(clouds, CFG_3.voxel_size, CFG_3.point_cloud_range, T, M, mean_features=5, padded=True, scatter=True)
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        cap = hard_voxelize_batch(clouds, CFG_3.voxel_size, CFG_3.point_cloud_range, T, M, mean_features=5, padded=True,
                                  scatter=True)
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(cap[0], voxels) and torch.equal(cap[5], canvas) and torch.equal(cap[3], base)


def test_canvas_without_a_mean_output_takes_the_fill_and_scatter_path():
    # C ABI with voxel_mean == NULL and a canvas: no mean rows to read the dense canvas pass from -> the canvas is
    # zero-filled on the side stream and the finalize kernel scatters into it; same bits as the dense pass
    import ctypes
    from mm_training_b200 import _lib
    from mm_training_b200.ops import voxelize as vzmod
    clouds = [torch.from_numpy(synthetic.lidar_sweep(n, 5, seed=40 + i)).to(DEV) for i, n in enumerate([30000, 52000])]
    T, M, mf = 15, 25000, 5
    ref = hard_voxelize_batch(clouds, CFG_3.voxel_size, CFG_3.point_cloud_range, T, M, mean_features=mf, padded=True, scatter=True)
    B, F = len(clouds), 5
    counts = [int(c.shape[0]) for c in clouds]
    offs = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    grid = vzmod._grid_size(CFG_3.voxel_size, CFG_3.point_cloud_range)
    gx, gy, gz = grid
    L = _lib.lib()
    tb = ctypes.c_size_t()
    _lib.check(L.bevvox_temp_bytes(B, int(offs[-1]), vzmod._i32_array(grid), M, T, ctypes.byref(tb)), 'bevvox_temp_bytes')
    pts = torch.cat(clouds, 0).contiguous()
    offsets = torch.from_numpy(offs).to(DEV)
    voxels = torch.empty(B * M, T, F, device=DEV)
    coors = torch.empty(B * M, 4, dtype=torch.int32, device=DEV)
    num = torch.empty(B * M, dtype=torch.int32, device=DEV)
    base = torch.empty(B + 1, dtype=torch.int32, device=DEV)
    canvas = torch.full((B, mf * gz, gy, gx), 7.0, device=DEV)             # (garbage: the call must fill it)
    temp = torch.empty(tb.value, dtype=torch.uint8, device=DEV)
    _lib.check(L.bevvox_hard_voxelize_scatter(
        pts.data_ptr(), None, offsets.data_ptr(), B, int(offs[-1]), max(counts), F, vzmod._f32_array(CFG_3.voxel_size),
        vzmod._f32_array(CFG_3.point_cloud_range), vzmod._i32_array(grid), T, M, voxels.data_ptr(), coors.data_ptr(),
        num.data_ptr(), base.data_ptr(), None, mf, canvas.data_ptr(), 0, temp.data_ptr(), _lib.stream_ptr(torch.device(DEV))),
        'bevvox_hard_voxelize_scatter')
    torch.cuda.synchronize()
    assert torch.equal(canvas, ref[5]) and torch.equal(voxels, ref[0]) and torch.equal(num, ref[1]) and torch.equal(base, ref[3])


@pytest.mark.parametrize('F,mf,T', [(12, 12, 5), (9, 7, 16), (16, 3, 2)])
def test_wide_point_rows_and_odd_mean_widths(F, mf, T):
    # point rows beyond 8 floats take the 16-wide instantiations of the finalize and canvas kernels; odd mean widths
    # split unevenly between the two half-warps that share a row's sums
    rng = np.random.default_rng(F * 100 + mf)
    clouds = []
    for n in (4000, 0, 2500):
        p = rng.uniform(-0.5, 8.5, size=(n, F)).astype(np.float32)
        p[:, 2] = rng.uniform(0.0, 1.0, n)
        clouds.append(p)
    vs, rg, M = [0.5, 0.5, 1.0], [0, 0, 0, 8, 8, 1], 200
    gpu = _check_batch(clouds, vs, rg, T, M, mean_features=mf)
    out = hard_voxelize_batch([torch.from_numpy(p).to(DEV) for p in clouds], vs, rg, T, M, mean_features=mf, padded=True,
                              scatter=True)
    n = int(out[3][-1])
    assert torch.equal(out[0][:n], gpu[0]) and torch.equal(out[4][:n], gpu[4])
    ref = vz.pillar_scatter_ref(gpu[4].cpu().numpy(), gpu[2].cpu().numpy(), 3, (1, 16, 16))
    assert np.array_equal(out[5].cpu().numpy(), ref)


def test_all_samples_empty():
    # no point at all: zero voxels, zero canvas, nothing launched over an empty range
    clouds = [torch.zeros(0, 5, device=DEV), torch.zeros(0, 5, device=DEV)]
    out = hard_voxelize_batch(clouds, CFG_3.voxel_size, CFG_3.point_cloud_range, 15, 100, mean_features=5, padded=True, scatter=True)
    assert out[3].tolist() == [0, 0, 0] and int(out[1].abs().sum()) == 0 and float(out[0].abs().sum()) == 0.0
    assert float(out[5].abs().sum()) == 0.0 and out[5].shape == (2, 5, 256, 2048)
    v, n, c, base, mean = hard_voxelize_batch(clouds, CFG_3.voxel_size, CFG_3.point_cloud_range, 15, 100, mean_features=5)
    assert v.shape[0] == 0 and n.shape[0] == 0 and c.shape == (0, 4)


def test_dynamic_voxelize():
    pts = synthetic.lidar_sweep(50000, 5, seed=3)
    coors = dynamic_voxelize(torch.from_numpy(pts).to(DEV), CFG_3.voxel_size, CFG_3.point_cloud_range)
    ref, _ = vz.point_coors_numpy(pts, CFG_3.voxel_size, CFG_3.point_cloud_range)
    assert np.array_equal(coors.cpu().numpy(), ref)
    layer = Voxelization(CFG_3.voxel_size, CFG_3.point_cloud_range, -1, -1)
    assert torch.equal(layer(torch.from_numpy(pts).to(DEV)), coors)


def test_vfe_module_and_train_eval_max_voxels():
    pts = torch.from_numpy(synthetic.lidar_sweep(20000, 8, seed=4)).to(DEV)
    layer = Voxelization(CFG_3.voxel_size, CFG_3.point_cloud_range, 15, (100, 200))
    layer.train()
    assert layer(pts)[0].shape[0] == 100
    layer.eval()
    voxels, num, coors, mean = voxelize([pts, pts.flip(0)], layer, mean_features=5)
    assert voxels.shape[0] == 400
    feats = HardSimpleVFE(5)(voxels, num, coors)
    assert feats.shape == (400, 5) and torch.allclose(feats, mean, rtol=1e-6, atol=1e-7)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        layer(pts.cpu())


@pytest.mark.parametrize('dtype', [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize('nz,C', [(1, 5), (1, 64), (3, 7)])
def test_pillar_scatter_forward_backward(dtype, nz, C):
    rng = np.random.default_rng(0)
    B, ny, nx = 3, 40, 96
    cells = rng.choice(B * nz * ny * nx, size=3000, replace=False)
    b, rem = np.divmod(cells, nz * ny * nx)
    z, rem = np.divmod(rem, ny * nx)
    y, x = np.divmod(rem, nx)
    coors = np.stack([b, z, y, x], 1).astype(np.int32)
    feats = torch.randn(3000, C).to(dtype)
    f = feats.to(DEV).requires_grad_(True)
    canvas = pillar_scatter(f, torch.from_numpy(coors).to(DEV), B, (nz, ny, nx))
    ref = vz.pillar_scatter_ref(feats.float().numpy(), coors, B, (nz, ny, nx))
    assert canvas.shape == (B, C * nz, ny, nx) and canvas.dtype == dtype
    assert np.array_equal(canvas.detach().float().cpu().numpy(), ref)                        # a copy: exact in any dtype
    g = torch.randn(B, C * nz, ny, nx).to(dtype)
    canvas.backward(g.to(DEV))
    gref = vz.pillar_scatter_backward_ref(g.float().numpy(), coors, (nz, ny, nx))
    assert np.array_equal(f.grad.float().cpu().numpy(), gref)


def test_lidar_branch_end_to_end_shapes():
    # models/bev_depth.py:181-183 with the shipped config: voxelize -> VFE -> scatter to (B, 5, 256, 2048)
    clouds = [torch.from_numpy(synthetic.lidar_sweep(100000, 8, seed=20 + i)).to(DEV) for i in range(2)]
    layer = Voxelization(CFG_3.voxel_size, CFG_3.point_cloud_range, 15, (25000, 25000)).eval()
    voxels, num_points, coors = voxelize(clouds, layer)
    feats = HardSimpleVFE(5)(voxels, num_points, coors)
    bev = PointPillarsScatter(5, (256, 2048))(feats, coors, 2)
    assert bev.shape == (2, 5, 256, 2048)
    rv, rn, rc = vz.voxelize_batch_ref([c.cpu().numpy() for c in clouds], CFG_3.voxel_size,
                                       CFG_3.point_cloud_range, 15, 25000)
    ref = vz.pillar_scatter_ref(vz.hard_simple_vfe_ref(rv, rn, 5), rc, 2, (1, 256, 2048))
    assert np.allclose(bev.cpu().numpy(), ref, rtol=1e-6, atol=1e-7)
    assert int((bev.abs().sum(1) > 0).sum()) <= voxels.shape[0]
