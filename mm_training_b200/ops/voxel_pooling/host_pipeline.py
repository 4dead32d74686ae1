"""Host-buffer entry point of the fused pooling path: pinned host tensors in, pinned host tensors out.

This is the call an integrator makes when the camera rig (or the reference's frustum geometry tensor), the depth
distributions and the context features live in host memory (data-loader side, or a C-ABI caller with host
buffers): the batch is cut into chunks of frames and pushed through three CUDA streams -- host->device copies,
compute (run plan + fused forward + fused backward through the public autograd op) and device->host copies -- so
that PCIe traffic in both directions overlaps with the kernels.  The op itself is unchanged; sharding by frame is
legal because no output element depends on another sample
(``ops/voxel_pooling/src/voxel_pooling_forward_cuda.cu:19`` of the reference).

Geometry comes in one of two forms:

* ``rig=(lsg, h_sensor2ego, h_intrin)`` -- what the data loader actually holds (``lss_fpn.py:328-361`` inputs): the
  two (B, N, 4, 4) matrices are copied (4 KB per frame instead of 3.8 MB of int32 indices) and the run plan is built
  on the device from ``sensor2ego @ inverse(intrin)`` (``rig.py``; indices bit-identical to the reference's ops);
* ``h_geom`` -- the reference's int32 (B, N, D, H, W, 3) tensor, copied chunk by chunk.

Run plans need a scratch row per run.  The run count of a chunk is read back once (first call: one sync per chunk)
and reused with 12 % headroom; every plan's status word is copied back with the outputs and checked after the
call -- a chunk whose geometry produced more runs than the scratch holds is detected (never silently wrong) and
the call is repeated with exact counts.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch

from .voxel_pooling import PoolingPlan, build_plan, runs_supported, voxel_pooling_fused


class HostPoolingPipeline:
    """Reusable staging buffers + streams for batches of frames of a fixed shape."""

    def __init__(self, num_cams: int, geom_shape: Sequence[int], depth_shape: Sequence[int],
                 context_shape: Sequence[int], voxel_num: Sequence[int], chunk_frames: int = 8,
                 dtype=torch.float32, device='cuda', rig=None):
        self.N = num_cams
        self.vn = tuple(int(v) for v in voxel_num)
        self.chunk = chunk_frames
        self.dev = torch.device(device)
        self.lsg = rig                    # LiftSplatGeometry or None
        self.variant = None
        if rig is not None:
            from .rig import rig_variant
            self.variant = rig_variant(self.dev)
        X, Y, _ = self.vn
        C = context_shape[1]
        self.frustum = (num_cams, *[int(v) for v in depth_shape[1:]])
        self.runs = runs_supported(C, dtype)
        n = chunk_frames
        mk = lambda shape, dt: [torch.empty(shape, dtype=dt, device=self.dev) for _ in range(2)]   # double buffers
        self.d_geom = mk((n, *geom_shape[1:]), torch.int32) if rig is None or self.variant is None else None
        self.d_depth = mk((n * num_cams, *depth_shape[1:]), dtype)
        self.d_ctx = mk((n * num_cams, *context_shape[1:]), dtype)
        self.d_go = mk((n, C, Y, X), dtype)
        self.s_in, self.s_run, self.s_out = (torch.cuda.Stream(self.dev) for _ in range(3))
        self.ev_in = [torch.cuda.Event() for _ in range(2)]
        self.ev_free = [torch.cuda.Event() for _ in range(2)]
        self.caps = {}                    # chunk index -> run_rows capacity (None until measured)
        self.h_status = None
        self.reruns = 0

    # -- plan of one chunk -------------------------------------------------------------------------------------
    def _plan(self, i: int, m: int, geom: Optional[torch.Tensor], combine: Optional[torch.Tensor]) -> PoolingPlan:
        cap = self.caps.get(i)
        if combine is not None:
            plan = PoolingPlan.from_rig(self.lsg, combine, self.variant, cap)
        else:
            plan = build_plan(geom, self.vn, frustum=self.frustum if self.runs else None, max_runs=cap)
        if plan.mode == 'runs' and cap is None:
            self.caps[i] = int(plan.num_sorted * 1.125) + 1024          # one read-back per chunk, first call only
        return plan

    def run(self, h_geom, h_depth, h_ctx, h_go, h_out, h_gdepth, h_gctx, h_sensor2ego=None, h_intrin=None,
            validate: bool = True):
        """All arguments are pinned host tensors; outputs are filled in place.  ``h_geom`` may be None when the
        pipeline was built with ``rig=`` and the two matrix tensors are given.  With ``validate`` (default) the call
        returns after the results are in host memory and the plans' status words are clean; ``validate=False`` only
        enqueues (call ``check()`` after synchronising)."""
        B, n, N = h_depth.shape[0] // self.N, self.chunk, self.N
        use_rig = self.lsg is not None and self.variant is not None and h_sensor2ego is not None
        if self.lsg is not None and not use_rig and h_geom is None:
            raise ValueError('no rig variant is proven on this device: pass h_geom (geometry.py / lsg.geom_xyz)')
        cur = torch.cuda.current_stream(self.dev)
        for s in (self.s_in, self.s_run, self.s_out):
            s.wait_stream(cur)
        nchunks = (B + n - 1) // n
        if self.h_status is None or self.h_status.numel() < nchunks:
            self.h_status = torch.zeros(nchunks, dtype=torch.int32).pin_memory()
        combine_all = None
        if use_rig:
            with torch.cuda.stream(self.s_in):
                d_s2e = h_sensor2ego.to(self.dev, non_blocking=True)
                d_k = h_intrin.to(self.dev, non_blocking=True)
                ev_rig = torch.cuda.Event()
                ev_rig.record(self.s_in)
            with torch.cuda.stream(self.s_run):
                self.s_run.wait_event(ev_rig)
                # lss_fpn.py:354 without the singularity check's host sync (same LU kernels, same values)
                combine_all = d_s2e.matmul(torch.linalg.inv_ex(d_k, check_errors=False).inverse).contiguous()
                d_s2e.record_stream(self.s_run)
                d_k.record_stream(self.s_run)
        for i, f0 in enumerate(range(0, B, n)):
            k, f1 = i & 1, min(f0 + n, B)
            m = f1 - f0
            with torch.cuda.stream(self.s_in):
                if i >= 2:
                    self.s_in.wait_event(self.ev_free[k])       # staging buffer k is free again
                if not use_rig:
                    self.d_geom[k][:m].copy_(h_geom[f0:f1], non_blocking=True)
                self.d_depth[k][:m * N].copy_(h_depth[f0 * N:f1 * N], non_blocking=True)
                self.d_ctx[k][:m * N].copy_(h_ctx[f0 * N:f1 * N], non_blocking=True)
                self.d_go[k][:m].copy_(h_go[f0:f1], non_blocking=True)
                self.ev_in[k].record(self.s_in)
            with torch.cuda.stream(self.s_run):
                self.s_run.wait_event(self.ev_in[k])
                d = self.d_depth[k][:m * N].detach().requires_grad_(True)
                c = self.d_ctx[k][:m * N].detach().requires_grad_(True)
                plan = self._plan(i, m, None if use_rig else self.d_geom[k][:m],
                                  combine_all[f0:f1] if use_rig else None)
                o = voxel_pooling_fused(None, d, c, self.vn, plan)
                o.backward(self.d_go[k][:m])
                self.ev_free[k].record(self.s_run)
                done = torch.cuda.Event()
                done.record(self.s_run)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(done)
                for t in (o, d.grad, c.grad, plan.buffer):
                    t.record_stream(self.s_out)
                h_out[f0:f1].copy_(o.detach(), non_blocking=True)
                h_gdepth[f0 * N:f1 * N].copy_(d.grad, non_blocking=True)
                h_gctx[f0 * N:f1 * N].copy_(c.grad, non_blocking=True)
                self.h_status[i:i + 1].copy_(plan.buffer.view(torch.int32)[2:3], non_blocking=True)   # PlanHeader.status
        for s in (self.s_in, self.s_run, self.s_out):
            cur.wait_stream(s)
        self._nchunks = nchunks
        if validate:
            cur.synchronize()
            if not self.check(raise_on_overflow=False):
                self.caps.clear()                               # measure the run counts of this geometry again
                self.reruns += 1
                self.run(h_geom, h_depth, h_ctx, h_go, h_out, h_gdepth, h_gctx, h_sensor2ego, h_intrin, validate=False)
                cur.synchronize()
                self.check()

    def check(self, raise_on_overflow: bool = True) -> bool:
        """After a synchronise: True when no chunk of the last call overflowed its run-row scratch."""
        ok = bool((self.h_status[:self._nchunks] == 0).all())
        if not ok and raise_on_overflow:
            raise RuntimeError('HostPoolingPipeline: run-row scratch overflow (geometry changed between calls); outputs invalid')
        return ok
