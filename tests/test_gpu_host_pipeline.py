"""GPU tests of the host-buffer entry point (ops/voxel_pooling/host_pipeline.py): pinned host tensors in and out,
run plans per chunk (from the camera rig or from the reference's geom_xyz tensor), scratch-overflow detection when
the geometry changes between calls.  Checked against the device-resident op (bit-equal: same kernels) and the
fp64 oracle (rtol 1e-5 + a few ulps of the summed magnitudes)."""
import pytest
import torch

from mm_training_b200 import synthetic
from mm_training_b200.configs import CFG_2
from mm_training_b200.ops.voxel_pooling import LiftSplatGeometry, build_plan, fused_backward, fused_forward, rig_variant
from mm_training_b200.ops.voxel_pooling.host_pipeline import HostPoolingPipeline
from oracle import voxel_pool_ref as vp

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _host_inputs(cfg, B, seed, jitter=5.0):
    s2e, intrin = synthetic.camera_rig_mats(cfg, B, yaw_jitter_deg=jitter, seed=seed)
    geom, vn = synthetic.camera_rig(cfg, B, device=DEV, yaw_jitter_deg=jitter, seed=seed)     # same ops, on the GPU
    depth, ctx, go = synthetic.camera_features(cfg, B, seed=seed)
    pin = lambda t: t.cpu().contiguous().pin_memory()
    return pin(s2e), pin(intrin), pin(geom), pin(depth), pin(ctx), pin(go), tuple(int(v) for v in vn.tolist())


def _outputs(cfg, B, vn):
    X, Y, _ = vn
    h, w = cfg.feat_hw
    return (torch.empty(B, cfg.output_channels, Y, X).pin_memory(),
            torch.empty(B * cfg.num_cams, cfg.depth_bins, h, w).pin_memory(),
            torch.empty(B * cfg.num_cams, cfg.output_channels, h, w).pin_memory())


def _device_result(geom, depth, ctx, go, vn, chunk=2, N=CFG_2.num_cams):
    """The device-resident op on the same chunking (the order in which stage B combines the partial sums of a cell
    is a pure function of the plan, so equal chunks give bit-equal results)."""
    outs, gds, gcs = [], [], []
    for f0 in range(0, geom.shape[0], chunk):
        f1 = min(f0 + chunk, geom.shape[0])
        g, d, c, o = (geom[f0:f1].to(DEV), depth[f0 * N:f1 * N].to(DEV), ctx[f0 * N:f1 * N].to(DEV), go[f0:f1].to(DEV))
        plan = build_plan(g, vn, frustum=tuple(g.shape[1:5]))
        outs.append(fused_forward(plan, d, c).permute(0, 3, 1, 2).cpu())
        gd, gc = fused_backward(plan, o, d, c)
        gds.append(gd.cpu())
        gcs.append(gc.cpu())
    return torch.cat(outs), torch.cat(gds), torch.cat(gcs)


@pytest.mark.parametrize('use_rig', [True, False])
def test_host_pipeline_matches_device_path_and_oracle(use_rig):
    cfg, B = CFG_2, 5                                     # 5 frames in chunks of 2: a ragged last chunk
    if use_rig and rig_variant(DEV) is None:
        pytest.skip('no proven rig variant on this device')
    s2e, intrin, geom, depth, ctx, go, vn = _host_inputs(cfg, B, seed=4)
    lsg = LiftSplatGeometry.from_config(cfg, DEV) if use_rig else None
    pipe = HostPoolingPipeline(cfg.num_cams, geom.shape, depth.shape, ctx.shape, vn, chunk_frames=2, device=DEV, rig=lsg)
    h_out, h_gd, h_gc = _outputs(cfg, B, vn)
    for _ in range(2):                                    # second call: hinted (sync-free) plans
        h_out.fill_(float('nan'))
        if use_rig:
            pipe.run(None, depth, ctx, go, h_out, h_gd, h_gc, h_sensor2ego=s2e, h_intrin=intrin)
        else:
            pipe.run(geom, depth, ctx, go, h_out, h_gd, h_gc)
        assert pipe.reruns == 0
        out, gd, gc = _device_result(geom, depth, ctx, go, vn)
        assert torch.equal(h_out, out) and torch.equal(h_gd, gd) and torch.equal(h_gc, gc)
    g = geom.cpu()
    feats = vp.materialise_features_ref(depth, ctx, B, cfg.num_cams)
    ref64 = vp.voxel_pooling_ref(g, feats, vn, acc_dtype=torch.float64)
    abs64 = vp.voxel_pooling_ref(g, feats.abs(), vn, acc_dtype=torch.float64)
    err = (h_out.double() - ref64).abs()
    assert bool((err <= 1e-5 * ref64.abs() + 1e-6 * abs64 + 1e-30).all())


def test_host_pipeline_detects_changed_geometry_and_reruns():
    cfg, B = CFG_2, 4
    if rig_variant(DEV) is None:
        pytest.skip('no proven rig variant on this device')
    lsg = LiftSplatGeometry.from_config(cfg, DEV)
    s2e, intrin, geom, depth, ctx, go, vn = _host_inputs(cfg, B, seed=4, jitter=0.0)
    pipe = HostPoolingPipeline(cfg.num_cams, geom.shape, depth.shape, ctx.shape, vn, chunk_frames=2, device=DEV, rig=lsg)
    h_out, h_gd, h_gc = _outputs(cfg, B, vn)
    pipe.run(None, depth, ctx, go, h_out, h_gd, h_gc, h_sensor2ego=s2e, h_intrin=intrin)
    pipe.caps = {i: max(1, c // 3) for i, c in pipe.caps.items()}        # as if the first geometry had far fewer runs
    pipe.run(None, depth, ctx, go, h_out, h_gd, h_gc, h_sensor2ego=s2e, h_intrin=intrin)
    assert pipe.reruns == 1                                               # flagged, repeated with exact counts
    out, gd, gc = _device_result(geom, depth, ctx, go, vn)
    assert torch.equal(h_out, out) and torch.equal(h_gd, gd) and torch.equal(h_gc, gc)
    pipe.caps = {i: max(1, c // 3) for i, c in pipe.caps.items()}
    pipe.run(None, depth, ctx, go, h_out, h_gd, h_gc, h_sensor2ego=s2e, h_intrin=intrin, validate=False)
    torch.cuda.synchronize()
    with pytest.raises(RuntimeError):
        pipe.check()
