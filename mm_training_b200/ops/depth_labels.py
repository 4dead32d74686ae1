"""Depth labels for the depth loss, on libbevpool_sm100 (csrc/depth_labels.cu).

Replaces ``exps/mm_training_aim.py:115-162`` (``get_depth_labels`` / ``get_depth_image``) and ``:180-215``
(``get_downsampled_gt_depth``) of the reference: a python triple loop over batch x sweep x camera with ~25 ATen
kernels and a matrix inverse per camera becomes two kernel launches for the whole batch.  Same argument meaning
and output as the reference pair ``get_downsampled_gt_depth(get_depth_labels(...))``:

    depth_labels(pointclouds, extrinsics, intrinsics, bda_mats, image_hw, downsample_factor, d_bound, depth_channels)
        -> (B * S * C * h * w, depth_channels) float32 one-hot

Several LiDAR points in one pixel: the last point in cloud order wins, deterministically (the reference's CPU
behaviour; its CUDA ``index_put_`` picks an arbitrary one).  No CPU fallback: CPU tensors raise.
"""
from __future__ import annotations

import ctypes
from typing import Sequence, Tuple

import torch

from .. import _lib
from .voxelize import _pointer_table


class DepthLabelGenerator:
    """Reusable scratch (one 64-bit word per full-resolution pixel, left clean by every call) for a fixed image layout."""

    def __init__(self, image_hw: Tuple[int, int], downsample_factor: int, d_bound: Sequence[float], depth_channels: int):
        self.H, self.W = int(image_hw[0]), int(image_hw[1])
        self.ds = int(downsample_factor)
        if self.H % self.ds or self.W % self.ds:
            raise ValueError('image size must be a multiple of the downsample factor (the reference views it that way)')
        # python-float arithmetic like the reference (`self.dbound[0] - self.dbound[2]`), then float32 at the tensor op
        self.bin_offset = float(d_bound[0]) - float(d_bound[2])
        self.bin_step = float(d_bound[2])
        self.D = int(depth_channels)
        self._scratch = None
        self._clean = False

    def __call__(self, pointclouds: Sequence[torch.Tensor], extrinsics: torch.Tensor, intrinsics: torch.Tensor,
                 bda_mats: torch.Tensor, return_bins: bool = False, bda_inv: torch.Tensor = None):
        """pointclouds: list of B float32 (Np_b, F >= 3) CUDA tensors; extrinsics / intrinsics (B, S, C, 4, 4);
        bda_mats (B, 4, 4).  Returns the one-hot labels (B*S*C*h*w, D) (and the int32 bin per cell).  ``bda_inv``
        (B, 3, 3): a precomputed ``inverse(bda_mats[:, :3, :3])`` to use instead of inverting here."""
        _lib.require_cuda(extrinsics, intrinsics, bda_mats, *pointclouds)
        B = len(pointclouds)
        assert extrinsics.shape[0] == B and extrinsics.shape[-2:] == (4, 4) and intrinsics.shape == extrinsics.shape
        images_per_sample = int(extrinsics.numel() // (16 * B))
        dev = extrinsics.device
        Fdim = int(pointclouds[0].shape[1])
        for p in pointclouds:
            if p.dtype != torch.float32 or p.dim() != 2 or p.shape[1] != Fdim:
                raise TypeError('pointclouds must be float32 (Np, F) tensors with the same F')
        clouds = [p.contiguous() for p in pointclouds]
        counts = [int(p.shape[0]) for p in clouds]
        L = _lib.lib()
        images = B * images_per_sample
        with torch.cuda.device(dev):
            nbytes = ctypes.c_size_t()
            _lib.check(L.bevlabel_scratch_bytes(images, self.H, self.W, ctypes.byref(nbytes)), 'bevlabel_scratch_bytes')
            if self._scratch is None or self._scratch.numel() < nbytes.value or self._scratch.device != dev:
                self._scratch = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
                self._clean = False
            ptrs = _pointer_table(tuple(p.data_ptr() for p in clouds), dev)
            cnt = torch.tensor(counts, dtype=torch.int32).to(dev, non_blocking=True)
            # :126-127 -- inverse(bda_mat[:3,:3]); B tiny matrices, left in torch (same LU kernels as the reference)
            if bda_inv is None:
                bda_inv = torch.linalg.inv(bda_mats[:, :3, :3].float())
            bda_inv = bda_inv.to(dev).float().contiguous()
            ext = extrinsics.float().contiguous()
            intr = intrinsics.float().contiguous()
            cells = images * (self.H // self.ds) * (self.W // self.ds)
            labels = torch.empty(cells, self.D, dtype=torch.float32, device=dev)
            bins = torch.empty(cells, dtype=torch.int32, device=dev) if return_bins else None
            _lib.check(L.bevlabel_depth_labels(ptrs.data_ptr(), cnt.data_ptr(), Fdim, B, images_per_sample, max(counts),
                                               bda_inv.data_ptr(), ext.data_ptr(), intr.data_ptr(), self.H, self.W, self.ds,
                                               self.bin_offset, self.bin_step, self.D, labels.data_ptr(),
                                               bins.data_ptr() if bins is not None else None, self._scratch.data_ptr(),
                                               1 if self._clean else 0, _lib.stream_ptr(dev)), 'bevlabel_depth_labels')
            self._clean = True
        return (labels, bins) if return_bins else labels


@torch.no_grad()
def depth_labels(pointclouds, extrinsics, intrinsics, bda_mats, image_hw, downsample_factor, d_bound, depth_channels,
                 return_bins: bool = False, bda_inv: torch.Tensor = None):
    """One-shot form of ``DepthLabelGenerator`` (allocates the scratch per call)."""
    return DepthLabelGenerator(image_hw, downsample_factor, d_bound, depth_channels)(
        pointclouds, extrinsics, intrinsics, bda_mats, return_bins, bda_inv)
