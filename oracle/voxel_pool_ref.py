"""CPU oracle for BEVDepth voxel pooling.  TEST INFRASTRUCTURE -- never imported by the product.

Pure-torch restatement of the reference operator.  Reference lines followed
(paths relative to the reference repo):

* bounds test, z-collapse and output address:
  ``ops/voxel_pooling/src/voxel_pooling_forward_cuda.cu:19-34``
* flattening, output allocation, ``pos_memo`` and the permuted return view:
  ``ops/voxel_pooling/voxel_pooling.py:30-40,55``
* backward gather: ``ops/voxel_pooling/voxel_pooling.py:58-69``
* depth (x) context outer product and layout fed to the op:
  ``layers/backbones/lss_fpn.py:441-460``

Pinned by ``test/test_ops/test_voxel_pooling.py:15-37`` (python-loop golden, re-run
in ``tests/test_oracle_voxel_pool.py``).
"""
from __future__ import annotations

import torch


def _voxel_num_ints(voxel_num):
    """``voxel_num`` is [X, Y, Z] (``voxel_pooling.py:37-38,45-47``)."""
    if isinstance(voxel_num, torch.Tensor):
        voxel_num = voxel_num.tolist()
    x, y, z = (int(v) for v in voxel_num)
    return x, y, z


def cell_index_ref(geom_xyz: torch.Tensor, voxel_num):
    """Per-point keep mask, linear cell index and the reference's ``pos_memo``.

    Follows ``voxel_pooling_forward_cuda.cu:19-29``: a point is dropped when any of
    x, y, z is outside ``[0, num_voxel)``; z only gates, it never addresses.

    Returns ``kept`` (B, Np) bool, ``lin`` (B, Np) int64 = b*Y*X + y*X + x (garbage
    where not kept) and ``pos_memo`` (B, Np, 3) int32 = (b, y, x) or -1.
    """
    X, Y, Z = _voxel_num_ints(voxel_num)
    B = geom_xyz.shape[0]
    g = geom_xyz.reshape(B, -1, 3).to(torch.int64)
    gx, gy, gz = g[..., 0], g[..., 1], g[..., 2]
    kept = (gx >= 0) & (gx < X) & (gy >= 0) & (gy < Y) & (gz >= 0) & (gz < Z)
    b = torch.arange(B, dtype=torch.int64).view(B, 1).expand_as(gx)
    lin = b * (Y * X) + gy * X + gx
    pos_memo = torch.full((B, g.shape[1], 3), -1, dtype=torch.int32)
    pos_memo[..., 0] = torch.where(kept, b, -1).to(torch.int32)
    pos_memo[..., 1] = torch.where(kept, gy, -1).to(torch.int32)
    pos_memo[..., 2] = torch.where(kept, gx, -1).to(torch.int32)
    return kept, lin, pos_memo


def voxel_pooling_ref(geom_xyz, input_features, voxel_num, acc_dtype=None):
    """Forward of ``voxel_pooling(geom_xyz, input_features, voxel_num)``.

    ``index_add_`` of kept rows into a zero (B*Y*X, C) grid, viewed (B, Y, X, C)
    and returned as the permuted (B, C, Y, X) view like ``voxel_pooling.py:55``.
    ``acc_dtype=torch.float64`` gives the error-bound twin.
    """
    X, Y, Z = _voxel_num_ints(voxel_num)
    B = input_features.shape[0]
    C = input_features.shape[-1]
    f = input_features.reshape(B, -1, C)
    kept, lin, _ = cell_index_ref(geom_xyz, voxel_num)
    assert kept.shape[1] == f.shape[1]
    dt = acc_dtype or input_features.dtype
    out = torch.zeros(B * Y * X, C, dtype=dt)
    k = kept.reshape(-1)
    out.index_add_(0, lin.reshape(-1)[k], f.reshape(-1, C)[k].to(dt))
    return out.view(B, Y, X, C).permute(0, 3, 1, 2)


def voxel_pooling_backward_ref(geom_xyz, grad_output, voxel_num, input_shape):
    """Backward of the op (``voxel_pooling.py:58-69``): kept rows receive
    ``grad_output[b, :, y, x]``, dropped rows receive zero; shape of the caller's
    ``input_features``."""
    X, Y, Z = _voxel_num_ints(voxel_num)
    B = grad_output.shape[0]
    C = grad_output.shape[1]
    kept, lin, _ = cell_index_ref(geom_xyz, voxel_num)
    g_rows = grad_output.permute(0, 2, 3, 1).reshape(B * Y * X, C)
    grad_in = torch.zeros(kept.numel(), C, dtype=grad_output.dtype)
    k = kept.reshape(-1)
    grad_in[k] = g_rows[lin.reshape(-1)[k]]
    return grad_in.reshape(input_shape)


def materialise_features_ref(depth, context, batch_size, num_cams):
    """``feat[b,n,d,h,w,c] = depth[b*N+n,d,h,w] * context[b*N+n,c,h,w]`` laid out
    (B, N, D, H, W, C) contiguous -- ``lss_fpn.py:441-443,447-454,460,463``."""
    f = depth.unsqueeze(1) * context.unsqueeze(2)            # (BN, C, D, H, W)
    f = f.reshape(batch_size, num_cams, *f.shape[1:])          # (B, N, C, D, H, W)
    return f.permute(0, 1, 3, 4, 5, 2).contiguous()            # (B, N, D, H, W, C)


def voxel_pooling_fused_ref(geom_xyz, depth, context, voxel_num, acc_dtype=None):
    """Oracle of the fused entry point: the reference's a5 step followed by the op.

    geom_xyz (B, N, D, H, W, 3) int32; depth (B*N, D, H, W); context (B*N, C, H, W).
    """
    B, N = geom_xyz.shape[0], geom_xyz.shape[1]
    feats = materialise_features_ref(depth, context, B, N)
    return voxel_pooling_ref(geom_xyz, feats, voxel_num, acc_dtype=acc_dtype)


def voxel_pooling_fused_grads_ref(geom_xyz, depth, context, voxel_num, grad_output,
                                  dtype=torch.float64):
    """(grad_depth, grad_context) by autograd through the pure-torch composition,
    computed in ``dtype`` (fp64 by default)."""
    d = depth.detach().to(dtype).requires_grad_(True)
    c = context.detach().to(dtype).requires_grad_(True)
    B, N = geom_xyz.shape[0], geom_xyz.shape[1]
    X, Y, Z = _voxel_num_ints(voxel_num)
    C = c.shape[1]
    feats = materialise_features_ref(d, c, B, N).reshape(-1, C)
    kept, lin, _ = cell_index_ref(geom_xyz, voxel_num)
    k = kept.reshape(-1)
    out = torch.zeros(B * Y * X, C, dtype=dtype).index_add(0, lin.reshape(-1)[k], feats[k])
    out = out.view(B, Y, X, C).permute(0, 3, 1, 2)
    out.backward(grad_output.to(dtype))
    return d.grad, c.grad


def python_loop_golden(geom_xyz_float, features, voxel_num=(128, 128, 1)):
    """The reference unit test's own golden, verbatim semantics
    (``test/test_ops/test_voxel_pooling.py:21-31``): ``.int()`` truncation, bounds
    test, ``+=`` in point order.  O(B*Np) python loop: small cases only."""
    X, Y, Z = voxel_num
    B = geom_xyz_float.shape[0]
    C = features.shape[-1]
    g = geom_xyz_float.reshape(B, -1, 3)
    f = features.reshape(B, -1, C)
    out = features.new_zeros(B, Y, X, C)
    for i in range(B):
        for j in range(g.shape[1]):
            x = g[i, j, 0].int()
            y = g[i, j, 1].int()
            z = g[i, j, 2].int()
            if x < 0 or x >= X or y < 0 or y >= Y or z < 0 or z >= Z:
                continue
            out[i, y, x, :] += f[i, j, :]
    return out.permute(0, 3, 1, 2)


def reference_test_inputs():
    """Inputs of the reference's only known-answer test, same seeds and recipe
    (``test/test_ops/test_voxel_pooling.py:15-20``)."""
    import numpy as np
    np.random.seed(0)
    torch.manual_seed(0)
    geom_xyz = torch.rand([2, 6, 10, 10, 10, 3]) * 160 - 80
    geom_xyz[..., 2] /= 100
    geom_xyz = geom_xyz.reshape(2, -1, 3)
    features = torch.rand([2, 6, 10, 10, 10, 80]) - 0.5
    return geom_xyz, features


def run_plan_ref(geom_xyz: torch.Tensor, voxel_num, rows_per_block: int = 16):
    """Expected contents of a RUN plan (mm_training_b200/csrc/common.cuh) for ``geom_xyz`` (B, N, D, H, W, 3).

    Not a reference data structure: the reference sums point by point
    (``voxel_pooling_forward_cuda.cu:30-34``).  A run groups vertically adjacent points -- same image,
    depth bin and column, consecutive rows inside one block of ``rows_per_block`` rows -- that the
    reference's own index rule (``cell_index_ref``) sends to the same cell, so summing run by run adds
    exactly the terms the reference adds, in a different (fixed) order.

    Returns ``head`` (B, N, D, H, W) bool, ``run_code`` (B, Np) int32 (slot of the run for its first
    point, -2 for continuation points, -1 for dropped points), ``cell_start`` (B*Y*X + 1,) CSR over
    runs, ``sorted_ids`` (R,) global ids of first points ordered by (cell, id)."""
    X, Y, Z = _voxel_num_ints(voxel_num)
    B, N, D, H, W = geom_xyz.shape[:5]
    kept, lin, _ = cell_index_ref(geom_xyz, voxel_num)
    cell = torch.where(kept, lin, torch.full_like(lin, -1)).reshape(B, N, D, H, W)
    prev = torch.full_like(cell, -1)
    prev[:, :, :, 1:, :] = cell[:, :, :, :-1, :]
    first_row = (torch.arange(H) % rows_per_block == 0).view(1, 1, 1, H, 1)
    head = (cell >= 0) & (first_row | (prev != cell))
    gid = torch.arange(cell.numel()).view_as(cell)
    hid, hcell = gid[head], cell[head]
    order = torch.sort(hcell, stable=True).indices
    sorted_ids = hid[order]
    counts = torch.bincount(hcell, minlength=B * X * Y)
    cell_start = torch.zeros(B * X * Y + 1, dtype=torch.int64)
    cell_start[1:] = counts.cumsum(0)
    code = torch.where(cell.reshape(-1) >= 0, torch.tensor(-2), torch.tensor(-1)).to(torch.int32)
    code[sorted_ids] = torch.arange(sorted_ids.numel(), dtype=torch.int32)
    return head, code.view(B, -1), cell_start, sorted_ids
