"""Value tests of the host-side geometry mirror: the frustum / voxel buffers of
``layers/backbones/lss_fpn.py:278-326``, ``get_geometry`` :328-361 against an independent fp64 computation, and the
index quantisation :461-462 on the shipped bounds (truncation toward zero, the -204.79998779 quirk, kept fractions of
the synthetic rigs quoted in SURVEY.md 8d)."""
import math

import numpy as np
import torch

from mm_training_b200 import geometry, synthetic
from mm_training_b200.configs import CFG_2, CFG_AIM, sweep_grid_config


def test_voxel_buffers_of_the_shipped_config():
    vs, vc, vn = geometry.voxel_buffers(CFG_AIM.x_bound, CFG_AIM.y_bound, CFG_AIM.z_bound)       # conf_aim.py:43-45
    assert vs.tolist() == [np.float32(0.8), np.float32(0.8), 8.0]
    assert vn.dtype == torch.int64 and vn.tolist() == [512, 64, 1]                                 # lss_fpn.py:286-289
    assert torch.allclose(vc, torch.tensor([-204.4, -25.2, -1.0]))
    lower = vc - vs / 2.0                                                                           # lss_fpn.py:461
    # NOT the float32 of the bound: (-204.8 + 0.4) - 0.4 in float32
    assert float(lower[0]) == float(np.float32(np.float32(-204.8 + 0.4) - np.float32(np.float32(0.8) / 2)))
    assert abs(float(lower[0]) - (-204.79998779296875)) < 1e-12 and float(lower[0]) != float(np.float32(-204.8))
    for g in (128, 256, 512):
        assert geometry.voxel_buffers(*[getattr(sweep_grid_config(g), k) for k in ('x_bound', 'y_bound', 'z_bound')])[2].tolist() == [g, g, 1]


def test_frustum_shape_and_values():
    fr = geometry.create_frustum(CFG_2.final_dim, CFG_2.downsample_factor, CFG_2.d_bound)         # lss_fpn.py:308-326
    assert fr.shape == (112, 16, 44, 4) and fr.dtype == torch.float32
    assert fr[0, 0, :, 0].tolist() == torch.linspace(0, 703, 44).tolist()
    assert fr[0, :, 0, 1].tolist() == torch.linspace(0, 255, 16).tolist()
    assert fr[:, 0, 0, 2].tolist() == torch.arange(2.0, 58.0, 0.5).tolist() and bool((fr[..., 3] == 1).all())
    aim = geometry.create_frustum(CFG_AIM.final_dim, CFG_AIM.downsample_factor, CFG_AIM.d_bound)
    assert aim.shape == (409, 44, 80, 4)                                                           # torch.arange(2, 206.4, 0.5): 409 bins


def test_get_geometry_against_fp64():
    cfg = CFG_2
    s2e, intrin = synthetic.camera_rig_mats(cfg, 2, yaw_jitter_deg=5.0, seed=3)
    fr = geometry.create_frustum(cfg.final_dim, cfg.downsample_factor, cfg.d_bound)
    pts = geometry.get_geometry(fr, s2e, intrin)
    assert pts.shape == (2, 4, 112, 16, 44, 3)
    # independent fp64 computation, point by point: p_cam = d * K^-1 (u, v, 1); p_ego = R p_cam + t
    f64 = fr.double()
    u, v, d = f64[..., 0], f64[..., 1], f64[..., 2]
    for b in range(2):
        for n in range(4):
            K, T = intrin[b, n].double(), s2e[b, n].double()
            cam = torch.stack([(u - K[0, 2]) / K[0, 0] * d, (v - K[1, 2]) / K[1, 1] * d, d], -1)
            ego = cam @ T[:3, :3].T + T[:3, 3]
            assert torch.allclose(pts[b, n].double(), ego, rtol=0, atol=2e-4)                      # fp32 rounding at 60 m
    # the first camera looks along +x from (1.5, 0, 1.6): the centre ray at depth 10 m
    mid = pts[0, 0, 16, 8, 22]
    assert abs(float(mid[0]) - 11.5) < 0.5 and abs(float(mid[1])) < 1.5


def test_quantisation_truncates_toward_zero_on_the_shipped_bounds():
    vs, vc, _ = geometry.voxel_buffers(CFG_AIM.x_bound, CFG_AIM.y_bound, CFG_AIM.z_bound)
    pts = torch.tensor([[-204.8, -25.6, -5.0], [-205.5, 0.0, -12.9], [-204.0, 25.59, 2.99], [204.79, 25.6, 3.0],
                        [-205.6, -26.5, -13.1], [0.0, 0.0, 0.0]])
    q = geometry.quantise_geometry(pts, vc, vs)
    assert q.dtype == torch.int32
    # .int() truncates: one extra voxel width BELOW the lower bound folds into cell 0; z in (-13, 3) is z-cell 0
    # (row 2 shows the lower-bound quirk: x = -204.0 is one voxel above the nominal bound -204.8 but (-204.0 +
    # 204.79998779) / 0.8 = 0.99998 -> cell 0, not 1; row 5: 204.79998779 / 0.8 = 255.99998 -> 255)
    assert q.tolist() == [[0, 0, 0], [0, 32, 0], [0, 63, 0], [511, 64, 1], [-1, -1, -1], [255, 32, 0]]
    # the float32 lower bound sits 1.2e-5 above -204.8: a point exactly on the nominal bound has a (tiny) negative offset
    assert float((torch.tensor(-204.8) - (vc - vs / 2)[0])) < 0


def test_kept_fractions_of_the_synthetic_rigs():
    # CFG-2 kept 47.7 %, CFG-AIM 25.7 %, 128^2 grid 59.3 % (level rig, no jitter)
    for cfg, want in ((CFG_2, 0.477), (CFG_AIM, 0.257), (sweep_grid_config(128), 0.593)):
        geom, vn = synthetic.camera_rig(cfg, 1)
        X, Y, Z = vn.tolist()
        g = geom.view(-1, 3)
        kept = (g[:, 0] >= 0) & (g[:, 0] < X) & (g[:, 1] >= 0) & (g[:, 1] < Y) & (g[:, 2] >= 0) & (g[:, 2] < Z)
        assert abs(float(kept.float().mean()) - want) < 0.002, (cfg.name, float(kept.float().mean()))
        assert geom.shape[1:5] == (cfg.num_cams, cfg.depth_bins, *cfg.feat_hw) and math.prod(geom.shape[1:5]) == cfg.points_per_frame
