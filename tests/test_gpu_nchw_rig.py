"""GPU parity tests of the round-2 kernels: the fused forward / backward reading the reference's NCHW tensors
through TMA tensor maps (csrc/pool_runs.cu, csrc/pool_bwd2.cu), the TMA gradient-row pass (csrc/pool.cu), the
run plan built straight from the camera rig (csrc/plan.cu, ops/voxel_pooling/rig.py) and the scratch-overflow
guard.  Oracle: oracle/voxel_pool_ref.py (fp32 / fp64 restatement of lss_fpn.py:441-464 + voxel_pooling.py);
integer results bit-exact, fp32 features / gradients rtol 1e-5 (+ an absolute term of a few ulps of the summed
magnitudes, see tests/test_gpu_voxel_pool.py)."""
import pytest
import torch

from mm_training_b200 import synthetic
from mm_training_b200.configs import CFG_2, CFG_AIM, sweep_grid_config
from mm_training_b200.ops.voxel_pooling import (LiftSplatGeometry, PoolingPlan, build_plan, fused_backward,
                                                fused_forward, rig_variant, voxel_pooling_fused, voxel_pooling_rig)
from mm_training_b200.ops.voxel_pooling.voxel_pooling import _grad_rows_nhwc, _nchw_direct
from oracle import voxel_pool_ref as vp

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _case(seed, B, N, D, H, W, C, vn, coherent):
    """coherent: every row of an (image, bin, column) lands in the same cell (a level camera) except rows that are
    dropped (z out of range) and a few that stray into another cell -> the kernels' fast path with its masks;
    else fully random geometry (every row its own cell: the per-row path)."""
    g = torch.Generator().manual_seed(seed)
    X, Y, Z = vn
    if coherent:
        x = torch.randint(-1, X + 1, (B, N, D, 1, W), generator=g).expand(B, N, D, H, W).clone()
        y = torch.randint(-1, Y + 1, (B, N, D, 1, W), generator=g).expand(B, N, D, H, W).clone()
        z = (torch.rand(B, N, D, H, W, generator=g) < 0.7).long() - 1 + torch.randint(0, Z, (B, N, D, H, W), generator=g)
        z = torch.where(z < 0, torch.full_like(z, -1), z)
        stray = torch.rand(B, N, D, H, W, generator=g) < 0.03
        x = torch.where(stray, torch.randint(0, X, (B, N, D, H, W), generator=g), x)
    else:
        x = torch.randint(-2, X + 2, (B, N, D, H, W), generator=g)
        y = torch.randint(-2, Y + 2, (B, N, D, H, W), generator=g)
        z = torch.randint(0, Z + 1, (B, N, D, H, W), generator=g)
    geom = torch.stack([x, y, z], -1).int().contiguous()
    depth = torch.rand(B * N, D, H, W, generator=g).softmax(1)
    ctx = torch.rand(B * N, C, H, W, generator=g) - 0.5
    go = torch.rand(B, C, Y, X, generator=g)
    return geom, depth, ctx, go


def _check(geom, depth, ctx, go, vn, channels_last=False, expect_direct=True):
    d = depth.cuda().requires_grad_(True)
    c = ctx.cuda()
    if channels_last:
        c = c.contiguous(memory_format=torch.channels_last)
    else:
        assert _nchw_direct(c, d) == expect_direct    # the TMA tensor-map entry points are the ones under test
    c.requires_grad_(True)
    out = voxel_pooling_fused(geom.cuda(), d, c, vn)
    B, N = geom.shape[0], geom.shape[1]
    feats = vp.materialise_features_ref(depth, ctx, B, N)
    ref64 = vp.voxel_pooling_ref(geom, feats, vn, acc_dtype=torch.float64)
    abs64 = vp.voxel_pooling_ref(geom, feats.abs(), vn, acc_dtype=torch.float64)
    err = (out.detach().cpu().double() - ref64).abs()
    assert bool((err <= 1e-5 * ref64.abs() + 1e-6 * abs64 + 1e-30).all()), float(err.max())
    assert bool((out.detach().cpu()[abs64 == 0] == 0).all())
    out.backward(go.cuda())
    gd, gc = vp.voxel_pooling_fused_grads_ref(geom, depth, ctx, vn, go)          # fp64
    C, D = ctx.shape[1], depth.shape[1]
    atol_d = 1.2e-7 * C * float(go.abs().max() * ctx.abs().max())
    atol_c = 1.2e-7 * D * float(go.abs().max() * depth.abs().max())
    assert torch.allclose(d.grad.double().cpu(), gd, rtol=1e-5, atol=atol_d)
    assert c.grad.shape == ctx.shape
    assert torch.allclose(c.grad.double().cpu(), gc, rtol=1e-5, atol=atol_c)
    kept, _, _ = vp.cell_index_ref(geom, vn)
    assert float(d.grad.detach().cpu().view(-1)[~kept.view(-1)].abs().sum()) == 0.0               # dropped points: exact zeros
    return out.detach(), d.grad.detach(), c.grad.detach()


SHAPES = [(2, 2, 20, 16, 8, 80, (32, 16, 1)),       # one full 16-row block, two column tiles
          (1, 3, 37, 44, 12, 32, (40, 12, 1)),      # H = 44 like the shipped config: ragged last row block, ragged last chunk
          (2, 1, 16, 5, 4, 128, (8, 8, 2)),         # H < 16, z gate with Z = 2
          (1, 2, 33, 17, 16, 64, (64, 8, 1)),       # H = 17: a one-row block
          (1, 1, 48, 16, 4, 96, (16, 16, 1))]


@pytest.mark.parametrize('shape', SHAPES)
@pytest.mark.parametrize('coherent', [True, False])
def test_nchw_tensor_map_forward_backward(shape, coherent):
    B, N, D, H, W, C, vn = shape
    _check(*_case(21, B, N, D, H, W, C, vn, coherent), vn)


@pytest.mark.parametrize('shape', SHAPES[:3])
def test_column_backward_on_pixel_rows(shape):
    B, N, D, H, W, C, vn = shape
    case = _case(22, B, N, D, H, W, C, vn, True)
    a = _check(*case, vn, channels_last=True)
    b = _check(*case, vn, channels_last=False)
    for x, y in zip(a, b):                           # same arithmetic, two layouts: same bits
        assert torch.equal(x.contiguous(), y.contiguous())


def test_nchw_path_on_camera_rigs_is_bit_stable_and_matches_rows_path(monkeypatch):
    for cfg, B in ((CFG_2, 2), (sweep_grid_config(256), 1)):
        geom, vn = synthetic.camera_rig(cfg, B, yaw_jitter_deg=5.0)
        depth, ctx, go = synthetic.camera_features(cfg, B)
        out, gd, gc = _check(geom, depth, ctx, go, vn.tolist())
        out2, gd2, gc2 = _check(geom, depth, ctx, go, vn.tolist())
        assert torch.equal(out, out2) and torch.equal(gd, gd2) and torch.equal(gc, gc2)
        monkeypatch.setenv('BEVPOOL_NCHW_DIRECT', '0')        # round-1 route: transposes + pixel-row kernels
        out3, gd3, gc3 = _check(geom, depth, ctx, go, vn.tolist(), expect_direct=False)
        monkeypatch.delenv('BEVPOOL_NCHW_DIRECT')
        assert torch.equal(out, out3) and torch.equal(gd, gd3) and torch.equal(gc, gc3)


def test_full_size_aim_against_oracle():
    # the shipped aiMotive shape (2 cams, D = 409, 44 x 80, P = 2.88 M points/frame), one frame, against the fp64 oracle
    cfg = CFG_AIM
    geom, vn = synthetic.camera_rig(cfg, 1)
    depth, ctx, go = synthetic.camera_features(cfg, 1)
    _check(geom, depth, ctx, go, vn.tolist())


@pytest.mark.parametrize('C,vn', [(80, (512, 64, 1)), (32, (100, 7, 1)), (128, (36, 5, 1)), (64, (128, 128, 1))])
def test_gradient_rows_tma(C, vn):
    X, Y, Z = vn
    B = 3
    g = torch.Generator().manual_seed(5)
    geom = torch.stack([torch.randint(-X // 2, X + X // 2, (B, 4000), generator=g),
                        torch.randint(0, Y, (B, 4000), generator=g), torch.zeros(B, 4000, dtype=torch.long)], -1).int()
    plan = build_plan(geom.cuda(), vn)
    go = torch.rand(B, C, Y, X, generator=g).cuda()
    rows = _grad_rows_nhwc(go, plan)
    cs = plan.cell_start.long()
    occ = (cs[1:] > cs[:-1]).view(B, Y, X)
    assert torch.equal(rows[occ], go.permute(0, 2, 3, 1)[occ])          # every occupied cell: its C gradients, bit-exact


# ---------------------------------------------------------------- plan from the rig (no geom_xyz tensor)
def test_rig_variant_is_proven_on_this_device():
    v = rig_variant(DEV)
    print('rig variant proven on this device:', v)
    assert v is not None, 'no accumulation order of the plan kernel reproduces torch on this device (fallback would be used)'


@pytest.mark.parametrize('cfg,B', [(CFG_2, 3), (CFG_AIM, 1), (sweep_grid_config(128), 2)])
def test_rig_plan_equals_geom_plan(cfg, B):
    v = rig_variant(DEV)
    if v is None:
        pytest.skip('no proven variant on this device')
    lsg = LiftSplatGeometry.from_config(cfg, DEV)
    gen = torch.Generator().manual_seed(3)
    s2e = torch.stack([torch.stack([synthetic.cam2ego(y + float(torch.rand(1, generator=gen)) * 10 - 5) for y in cfg.cam_yaws_deg])
                       for _ in range(B)]).to(DEV)
    k = torch.eye(4)
    k[0, 0] = k[1, 1] = cfg.focal_px
    k[0, 2], k[1, 2] = cfg.final_dim[1] / 2, cfg.final_dim[0] / 2
    intrin = k[None, None].repeat(B, cfg.num_cams, 1, 1).to(DEV)
    geom = lsg.geom_xyz(s2e, intrin)                                      # the reference's ops (geometry.py)
    assert torch.equal(lsg.rig_geom(lsg.combine(s2e, intrin), v), geom)  # bit-exact integer indices
    ref = PoolingPlan(geom, lsg.voxel_num, frustum=tuple(geom.shape[1:5]))
    rig = lsg.plan(s2e, intrin)
    assert rig.mode == ref.mode == 'runs'
    assert torch.equal(rig.cell_of_point, ref.cell_of_point)
    assert torch.equal(rig.cell_start, ref.cell_start)
    assert torch.equal(rig.sorted_ids, ref.sorted_ids)
    assert torch.equal(rig.sorted_cells, ref.sorted_cells)
    assert torch.equal(rig.run_code, ref.run_code)
    depth, ctx, go = synthetic.camera_features(cfg, B, device=DEV)
    a = voxel_pooling_rig(lsg, s2e, intrin, depth, ctx)
    b = voxel_pooling_fused(None, depth, ctx, lsg.voxel_num, ref)
    assert torch.equal(a, b)


def test_rig_plan_with_tilted_cameras_matches_oracle():
    v = rig_variant(DEV)
    if v is None:
        pytest.skip('no proven variant on this device')
    from mm_training_b200.ops.voxel_pooling.rig import _random_rigs
    cfg = CFG_2
    lsg = LiftSplatGeometry.from_config(cfg, DEV)
    s2e, k = _random_rigs(2, cfg.num_cams, torch.Generator().manual_seed(9))
    s2e, k = s2e.to(DEV), k.to(DEV)
    geom = lsg.geom_xyz(s2e, k)
    plan = lsg.plan(s2e, k)
    assert torch.equal(plan.cell_of_point.view(-1), build_plan(geom, lsg.voxel_num).cell_of_point.view(-1))
    depth, ctx, go = synthetic.camera_features(cfg, 2)
    d, c = depth.cuda().requires_grad_(True), ctx.cuda().requires_grad_(True)
    out = voxel_pooling_fused(None, d, c, lsg.voxel_num, plan)
    out.backward(go.cuda())
    g = geom.cpu()
    feats = vp.materialise_features_ref(depth, ctx, 2, cfg.num_cams)
    ref64 = vp.voxel_pooling_ref(g, feats, lsg.voxel_num, acc_dtype=torch.float64)
    abs64 = vp.voxel_pooling_ref(g, feats.abs(), lsg.voxel_num, acc_dtype=torch.float64)
    err = (out.detach().cpu().double() - ref64).abs()
    assert bool((err <= 1e-5 * ref64.abs() + 1e-6 * abs64 + 1e-30).all())
    gd, gc = vp.voxel_pooling_fused_grads_ref(g, depth, ctx, lsg.voxel_num, go)
    assert torch.allclose(d.grad.double().cpu(), gd, rtol=1e-5, atol=1.2e-7 * 80 * 0.5)
    assert torch.allclose(c.grad.double().cpu(), gc, rtol=1e-5, atol=1.2e-7 * cfg.depth_bins)


# ---------------------------------------------------------------- scratch-row overflow guard
def test_stale_max_runs_hint_is_flagged_not_fatal():
    cfg = CFG_2
    geom, vn = synthetic.camera_rig(cfg, 2, device=DEV, yaw_jitter_deg=5.0)
    depth, ctx, _ = synthetic.camera_features(cfg, 2, device=DEV)
    fr = tuple(geom.shape[1:5])
    exact = build_plan(geom, vn, frustum=fr)
    n = exact.num_sorted
    good = fused_forward(exact, depth, ctx)
    assert exact.status() == 0
    ok = build_plan(geom, vn, frustum=fr, max_runs=n)
    assert torch.equal(fused_forward(ok, depth, ctx), good) and ok.status() == 0
    stale = build_plan(geom, vn, frustum=fr, max_runs=n // 2)              # a hint from some other geometry
    fused_forward(stale, depth, ctx)                                        # must not touch memory beyond the scratch
    torch.cuda.synchronize()
    assert stale.status() == 1
    with pytest.raises(RuntimeError):
        stale.raise_if_overflowed()


# ---------------------------------------------------------------- concat epilogue (bev_depth.py:187-189)
@pytest.mark.parametrize('channels_last_ctx', [False, True])
@pytest.mark.parametrize('c2', [256, 4])
def test_concat_epilogue_equals_cat_of_the_stock_op(channels_last_ctx, c2):
    from mm_training_b200.ops.voxel_pooling import voxel_pooling_fused_concat
    cfg, B = CFG_2, 3
    geom, vn_t = synthetic.camera_rig(cfg, B, device=DEV, yaw_jitter_deg=5.0, seed=8)
    vn = tuple(int(v) for v in vn_t.tolist())
    depth, ctx, _ = synthetic.camera_features(cfg, B, device=DEV, seed=8)
    if channels_last_ctx:
        ctx = ctx.contiguous(memory_format=torch.channels_last)
    X, Y, _ = vn
    g = torch.Generator().manual_seed(2)
    other = torch.randn(B, c2, Y, X, generator=g).to(DEV)
    go = torch.rand(B, cfg.output_channels + c2, Y, X, generator=g).to(DEV).contiguous(memory_format=torch.channels_last)
    plan = build_plan(geom, vn, frustum=tuple(geom.shape[1:5]))
    d1, c1, o1 = depth.clone().requires_grad_(True), ctx.clone().requires_grad_(True), other.clone().requires_grad_(True)
    d2, c2_, o2 = depth.clone().requires_grad_(True), ctx.clone().requires_grad_(True), other.clone().requires_grad_(True)
    cat = voxel_pooling_fused_concat(d1, c1, o1, vn, plan)
    ref = torch.cat([voxel_pooling_fused(None, d2, c2_, vn, plan), o2], dim=1)
    assert cat.shape == ref.shape and cat.permute(0, 2, 3, 1).is_contiguous()
    assert torch.equal(cat, ref)                                       # same kernels, same order: bit-equal
    cat.backward(go)
    ref.backward(go)
    assert torch.equal(d1.grad, d2.grad) and torch.equal(c1.grad, c2_.grad) and torch.equal(o1.grad, o2.grad)
    go_nchw = go.contiguous()                                           # an NCHW gradient takes the copying path: same values
    d1.grad = c1.grad = o1.grad = None
    voxel_pooling_fused_concat(d1, c1, o1, vn, plan).backward(go_nchw)
    assert torch.equal(d1.grad, d2.grad) and torch.equal(c1.grad, c2_.grad) and torch.equal(o1.grad, o2.grad)
