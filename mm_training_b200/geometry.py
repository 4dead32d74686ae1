"""Host-side mirror of the reference's frustum geometry and cell quantisation.

These stay in PyTorch on purpose: cell indices must be bit-identical to what the
reference computes, and the reference computes them with exactly these ATen ops
(``layers/backbones/lss_fpn.py``):

* buffers ``voxel_size / voxel_coord / voxel_num``      -- ``lss_fpn.py:278-289``
* ``create_frustum``                                    -- ``lss_fpn.py:308-326``
* ``get_geometry``                                      -- ``lss_fpn.py:328-361``
* quantisation to int32 cells (truncation toward zero)  -- ``lss_fpn.py:461-462``
"""
from __future__ import annotations

import torch


def voxel_buffers(x_bound, y_bound, z_bound):
    """(voxel_size f32[3], voxel_coord f32[3], voxel_num int64[3]) as registered at
    ``lss_fpn.py:278-289``."""
    rows = [x_bound, y_bound, z_bound]
    voxel_size = torch.Tensor([row[2] for row in rows])
    voxel_coord = torch.Tensor([row[0] + row[2] / 2.0 for row in rows])
    voxel_num = torch.LongTensor([(row[1] - row[0]) / row[2] for row in rows])
    return voxel_size, voxel_coord, voxel_num


def create_frustum(final_dim, downsample_factor, d_bound):
    """(D, H, W, 4) image-plane grid (u, v, d, 1) -- ``lss_fpn.py:308-326``."""
    ogfH, ogfW = final_dim
    fH, fW = ogfH // downsample_factor, ogfW // downsample_factor
    d_coords = torch.arange(*d_bound, dtype=torch.float).view(-1, 1, 1).expand(-1, fH, fW)
    D = d_coords.shape[0]
    x_coords = torch.linspace(0, ogfW - 1, fW, dtype=torch.float).view(1, 1, fW).expand(D, fH, fW)
    y_coords = torch.linspace(0, ogfH - 1, fH, dtype=torch.float).view(1, fH, 1).expand(D, fH, fW)
    paddings = torch.ones_like(d_coords)
    return torch.stack((x_coords, y_coords, d_coords, paddings), -1)


def get_geometry(frustum, sensor2ego_mat, intrin_mat):
    """Camera frustum -> ego xyz, (B, N, D, H, W, 3) float32 -- ``lss_fpn.py:328-361``
    (the BDA branch is commented out in the reference; BDA is applied later on the
    BEV map, ``models/bev_depth.py:176``)."""
    batch_size, num_cams, _, _ = sensor2ego_mat.shape
    points = frustum.repeat(batch_size, num_cams, 1, 1, 1, 1).unsqueeze(-1)
    points = torch.cat((points[:, :, :, :, :, :2] * points[:, :, :, :, :, 2:3],
                        points[:, :, :, :, :, 2:]), 5)
    combine = sensor2ego_mat.matmul(torch.inverse(intrin_mat))
    points = combine.view(batch_size, num_cams, 1, 1, 1, 4, 4).matmul(points)
    points = points.squeeze(-1)
    return points[..., :3]


def quantise_geometry(geom_xyz, voxel_coord, voxel_size):
    """float ego xyz -> int32 cell index, truncation toward zero -- ``lss_fpn.py:461-462``."""
    return ((geom_xyz - (voxel_coord - voxel_size / 2.0)) / voxel_size).int()
