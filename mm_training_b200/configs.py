"""Shapes of the BEV-projection path, named after the reference's config fields
(``exps/conf_aim.py`` of the reference; numbers derived in SURVEY.md section 8)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Sequence, Tuple


@dataclass(frozen=True)
class CameraPoolConfig:
    """Camera frustum -> BEV grid pooling configuration (``backbone_conf`` fields)."""
    name: str
    final_dim: Tuple[int, int]             # (H_img, W_img)            conf_aim.py:3
    downsample_factor: int                 # conf_aim.py:51-52
    d_bound: Tuple[float, float, float]    # conf_aim.py:46
    x_bound: Tuple[float, float, float]    # conf_aim.py:43
    y_bound: Tuple[float, float, float]    # conf_aim.py:44
    z_bound: Tuple[float, float, float]    # conf_aim.py:45
    output_channels: int                   # conf_aim.py:36 (camera_feature_channels)
    focal_px: float                        # synthetic rig only
    cam_yaws_deg: Tuple[float, ...]        # synthetic rig only

    @property
    def num_cams(self) -> int:
        return len(self.cam_yaws_deg)

    @property
    def feat_hw(self) -> Tuple[int, int]:
        return (self.final_dim[0] // self.downsample_factor,
                self.final_dim[1] // self.downsample_factor)

    @property
    def depth_bins(self) -> int:
        import torch
        return int(torch.arange(*self.d_bound, dtype=torch.float).numel())

    @property
    def voxel_num(self) -> Tuple[int, int, int]:
        # same truncation as lss_fpn.py:286-289 (LongTensor of float quotients)
        return tuple(int((r[1] - r[0]) / r[2]) for r in (self.x_bound, self.y_bound, self.z_bound))

    @property
    def points_per_frame(self) -> int:
        h, w = self.feat_hw
        return self.num_cams * self.depth_bins * h * w


_AIM_BOUNDS = dict(x_bound=(-204.8, 204.8, 0.8), y_bound=(-25.6, 25.6, 0.8), z_bound=(-5.0, 3.0, 8.0))

#: BASELINE.json configs[1]: aiMotive 4-cam, D=112, 16x44 feature map, C=80 on the aiMotive BEV grid.
CFG_2 = CameraPoolConfig(name='cfg2_aim4cam_D112_16x44_C80_grid512x64', final_dim=(256, 704),
                         downsample_factor=16, d_bound=(2.0, 58.0, 0.5), output_channels=80,
                         focal_px=560.0, cam_yaws_deg=(0.0, 180.0, 90.0, -90.0), **_AIM_BOUNDS)

#: What exps/conf_aim.py actually ships: 2 pinhole views, 704x1280 / 16, d_bound [2, 206.4, 0.5] -> 409 bins.
CFG_AIM = CameraPoolConfig(name='aim_shipped_2cam_D409_44x80_C80_grid512x64', final_dim=(704, 1280),
                           downsample_factor=16, d_bound=(2.0, 206.4, 0.5), output_channels=80,
                           focal_px=1250.0, cam_yaws_deg=(0.0, 180.0), **_AIM_BOUNDS)


def sweep_grid_config(grid: int) -> CameraPoolConfig:
    """BASELINE.json configs[4]: square grids 128/256/512 over +-51.2 m."""
    step = {128: 0.8, 256: 0.4, 512: 0.2}[grid]
    return CameraPoolConfig(name=f'cfg5_4cam_D112_16x44_C80_grid{grid}x{grid}', final_dim=(256, 704),
                            downsample_factor=16, d_bound=(2.0, 58.0, 0.5), output_channels=80,
                            focal_px=560.0, cam_yaws_deg=(0.0, 180.0, 90.0, -90.0),
                            x_bound=(-51.2, 51.2, step), y_bound=(-51.2, 51.2, step),
                            z_bound=(-5.0, 3.0, 8.0))


@dataclass(frozen=True)
class VoxelizerConfig:
    """``lidar_conf['pts_voxel_layer']`` (conf_aim.py:16-18,192-197)."""
    name: str = 'cfg3_lidar200k_F5_vs0.2_T15_M25000'
    voxel_size: Tuple[float, float, float] = (0.2, 0.2, 8.0)
    point_cloud_range: Tuple[float, ...] = (-204.8, -25.6, -5.0, 204.8, 25.6, 3.0)
    max_num_points: int = 15
    max_voxels: int = 25000
    num_point_features: int = 5            # 8 with radar (data_loader.py:324-330)
    vfe_features: int = 5                  # HardSimpleVFE num_features, conf_aim.py:198-201
    points_per_sweep: int = 200_000        # MAX_LIDAR_POINTS * sweeps, data_loader.py:71


CFG_3 = VoxelizerConfig()
