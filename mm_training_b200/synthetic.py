"""Deterministic synthetic inputs of the reference's shapes.

No dataset is needed: a pinhole camera rig produces the frustum -> cell indices through
the reference's own geometry ops (``mm_training_b200.geometry``), and a long-range
LiDAR sweep generator produces point clouds in the aiMotive column layout
(``dataset/src/data_loader.py:324-330`` of the reference).
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import geometry
from .configs import CameraPoolConfig, VoxelizerConfig


def cam2ego(yaw_deg: float, t=(1.5, 0.0, 1.6)) -> torch.Tensor:
    """4x4 camera->ego: Rz(yaw) * R0 with the reference's camera->body convention
    (``dataset/src/data_loader.py:37-39``)."""
    r0 = torch.tensor([[0, 0, 1.], [-1, 0, 0], [0, -1, 0]])
    c, s = math.cos(math.radians(yaw_deg)), math.sin(math.radians(yaw_deg))
    rz = torch.tensor([[c, -s, 0], [s, c, 0], [0, 0, 1.]])
    m = torch.eye(4)
    m[:3, :3] = rz @ r0
    m[:3, 3] = torch.tensor(t)
    return m


def camera_rig_mats(cfg: CameraPoolConfig, batch_size: int, device='cpu', yaw_jitter_deg: float = 0.0, seed: int = 1):
    """(sensor2ego_mat, intrin_mat), both (B, N, 4, 4) float32: what ``LSSFPN.forward`` receives in ``mats_dict``
    (``lss_fpn.py:469-474``).  ``yaw_jitter_deg`` > 0 draws a per-sample yaw offset U(-j, j) (the "cold plan" runs)."""
    k = torch.eye(4)
    k[0, 0] = cfg.focal_px
    k[1, 1] = cfg.focal_px
    k[0, 2] = cfg.final_dim[1] / 2
    k[1, 2] = cfg.final_dim[0] / 2
    n = cfg.num_cams
    gen = torch.Generator().manual_seed(seed)
    mats = []
    for _ in range(batch_size):
        j = (torch.rand(1, generator=gen).item() * 2 - 1) * yaw_jitter_deg if yaw_jitter_deg else 0.0
        mats.append(torch.stack([cam2ego(y + j) for y in cfg.cam_yaws_deg]))
    s2e = torch.stack(mats).to(device)
    intrin = k[None, None].repeat(batch_size, n, 1, 1).to(device)
    return s2e, intrin


def camera_rig(cfg: CameraPoolConfig, batch_size: int, device='cpu', yaw_jitter_deg: float = 0.0,
               seed: int = 1):
    """Returns (geom_xyz int32 (B,N,D,H,W,3), voxel_num int64[3]) for the rig of ``cfg``: the reference's own
    geometry ops (``mm_training_b200.geometry``) applied to ``camera_rig_mats``."""
    frustum = geometry.create_frustum(cfg.final_dim, cfg.downsample_factor, cfg.d_bound).to(device)
    s2e, intrin = camera_rig_mats(cfg, batch_size, device, yaw_jitter_deg, seed)
    voxel_size, voxel_coord, voxel_num = geometry.voxel_buffers(cfg.x_bound, cfg.y_bound, cfg.z_bound)
    pts = geometry.get_geometry(frustum, s2e, intrin)
    geom = geometry.quantise_geometry(pts, voxel_coord.to(device), voxel_size.to(device))
    return geom.contiguous(), voxel_num


def camera_features(cfg: CameraPoolConfig, batch_size: int, device='cpu', dtype=torch.float32,
                    seed: int = 1):
    """depth = softmax(rand) over D, context = rand - 0.5, grad_out = rand."""
    gen = torch.Generator(device='cpu').manual_seed(seed)
    h, w = cfg.feat_hw
    bn = batch_size * cfg.num_cams
    x, y, _ = cfg.voxel_num
    depth = torch.rand(bn, cfg.depth_bins, h, w, generator=gen).softmax(1)
    context = torch.rand(bn, cfg.output_channels, h, w, generator=gen) - 0.5
    grad_out = torch.rand(batch_size, cfg.output_channels, y, x, generator=gen)
    return (depth.to(device=device, dtype=dtype), context.to(device=device, dtype=dtype),
            grad_out.to(device=device, dtype=dtype))


def lidar_sweep(num_points: int = 200_000, num_features: int = 5, seed: int = 2,
                num_radar: int = 2000) -> np.ndarray:
    """Long-range synthetic sweep: r ~ 204.8 U^1.5,
    azimuth ~ U(-pi, pi), z ~ N(-0.5, 1.5) clipped to [-8, 6].  F=5: [x,y,z,intensity,t];
    F=8: [x,y,z,is_radar,speed,power,intensity,t] with the first ``num_radar`` rows radar."""
    rng = np.random.default_rng(seed)
    r = 204.8 * rng.random(num_points) ** 1.5
    az = rng.uniform(-np.pi, np.pi, num_points)
    x, y = r * np.cos(az), r * np.sin(az)
    z = np.clip(rng.normal(-0.5, 1.5, num_points), -8.0, 6.0)
    inten, t = rng.random(num_points), rng.random(num_points)
    if num_features == 5:
        cols = [x, y, z, inten, t]
    elif num_features == 8:
        is_radar = np.zeros(num_points)
        is_radar[:num_radar] = 1.0
        speed = rng.normal(0, 5, num_points) * is_radar
        power = rng.random(num_points) * is_radar
        cols = [x, y, z, is_radar, speed, power, inten, t]
    else:
        cols = [x, y, z] + [rng.random(num_points) for _ in range(num_features - 3)]
    return np.stack(cols, 1).astype(np.float32)
