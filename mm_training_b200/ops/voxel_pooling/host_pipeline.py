"""Host-buffer entry point of the fused pooling path: pinned host tensors in, pinned host tensors out.

This is the call an integrator makes when the frustum geometry, depth distributions and context
features live in host memory (data-loader side, or a C-ABI caller with host buffers): the batch is cut
into chunks of frames and pushed through three CUDA streams -- host->device copies, compute (plan build
+ fused forward + fused backward through the public autograd op) and device->host copies -- so that
PCIe traffic in both directions overlaps with the kernels.  The op itself is unchanged; sharding by
frame is legal because no output element depends on another sample
(``ops/voxel_pooling/src/voxel_pooling_forward_cuda.cu:19`` of the reference).
"""
from __future__ import annotations

from typing import Sequence

import torch

from .voxel_pooling import build_plan, voxel_pooling_fused


class HostPoolingPipeline:
    """Reusable staging buffers + streams for ``frames`` frames of a fixed shape."""

    def __init__(self, num_cams: int, geom_shape: Sequence[int], depth_shape: Sequence[int],
                 context_shape: Sequence[int], voxel_num: Sequence[int], chunk_frames: int = 8,
                 dtype=torch.float32, device='cuda'):
        self.N = num_cams
        self.vn = tuple(int(v) for v in voxel_num)
        self.chunk = chunk_frames
        self.dev = torch.device(device)
        X, Y, _ = self.vn
        C = context_shape[1]
        n = chunk_frames
        mk = lambda shape, dt: [torch.empty(shape, dtype=dt, device=self.dev) for _ in range(2)]   # double buffers
        self.d_geom = mk((n, *geom_shape[1:]), torch.int32)
        self.d_depth = mk((n * num_cams, *depth_shape[1:]), dtype)
        self.d_ctx = mk((n * num_cams, *context_shape[1:]), dtype)
        self.d_go = mk((n, C, Y, X), dtype)
        self.s_in, self.s_run, self.s_out = (torch.cuda.Stream(self.dev) for _ in range(3))
        self.ev_in = [torch.cuda.Event() for _ in range(2)]
        self.ev_free = [torch.cuda.Event() for _ in range(2)]

    def run(self, h_geom, h_depth, h_ctx, h_go, h_out, h_gdepth, h_gctx):
        """All arguments are pinned host tensors; outputs are filled in place.  Returns after enqueueing;
        call ``torch.cuda.synchronize()`` (or wait on the current stream) before reading the outputs."""
        B, n, N = h_geom.shape[0], self.chunk, self.N
        cur = torch.cuda.current_stream(self.dev)
        for s in (self.s_in, self.s_run, self.s_out):
            s.wait_stream(cur)
        for i, f0 in enumerate(range(0, B, n)):
            k, f1 = i & 1, min(f0 + n, B)
            m = f1 - f0
            with torch.cuda.stream(self.s_in):
                if i >= 2:
                    self.s_in.wait_event(self.ev_free[k])       # staging buffer k is free again
                self.d_geom[k][:m].copy_(h_geom[f0:f1], non_blocking=True)
                self.d_depth[k][:m * N].copy_(h_depth[f0 * N:f1 * N], non_blocking=True)
                self.d_ctx[k][:m * N].copy_(h_ctx[f0 * N:f1 * N], non_blocking=True)
                self.d_go[k][:m].copy_(h_go[f0:f1], non_blocking=True)
                self.ev_in[k].record(self.s_in)
            with torch.cuda.stream(self.s_run):
                self.s_run.wait_event(self.ev_in[k])
                d = self.d_depth[k][:m * N].detach().requires_grad_(True)
                c = self.d_ctx[k][:m * N].detach().requires_grad_(True)
                # point plan: sizing the scratch rows of a run plan would read the run count back (a host
                # sync per chunk that stalls the copy streams); a chunk's kernels hide behind PCIe anyway
                plan = build_plan(self.d_geom[k][:m], self.vn)
                o = voxel_pooling_fused(None, d, c, self.vn, plan)
                o.backward(self.d_go[k][:m])
                self.ev_free[k].record(self.s_run)
                done = torch.cuda.Event()
                done.record(self.s_run)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(done)
                for t in (o, d.grad, c.grad):
                    t.record_stream(self.s_out)
                h_out[f0:f1].copy_(o.detach(), non_blocking=True)
                h_gdepth[f0 * N:f1 * N].copy_(d.grad, non_blocking=True)
                h_gctx[f0 * N:f1 * N].copy_(c.grad, non_blocking=True)
        for s in (self.s_in, self.s_run, self.s_out):
            cur.wait_stream(s)
