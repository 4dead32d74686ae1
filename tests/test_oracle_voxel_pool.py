"""Pins the CPU oracle of voxel pooling to the reference's own known-answer test and to
the committed golden fixture (no GPU needed)."""
import hashlib
import os

import numpy as np
import torch

from oracle import voxel_pool_ref as vp

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')


def test_oracle_matches_reference_test_golden_loop():
    # reference recipe: test/test_ops/test_voxel_pooling.py:15-37 (rtol 1e-3 there; exact here)
    geom, feats = vp.reference_test_inputs()
    gold = vp.python_loop_golden(geom, feats, (128, 128, 1))
    out = vp.voxel_pooling_ref(geom.int(), feats, torch.tensor([128, 128, 1], dtype=torch.int))
    assert out.shape == (2, 80, 128, 128)
    assert torch.allclose(gold, out, rtol=1e-3)
    # index_add_ visits kept points in ascending order per cell, like the loop: bit-equal
    assert torch.equal(gold, out)


def test_oracle_matches_committed_fixture():
    fx = np.load(os.path.join(GOLDEN, 'voxel_pool_reftest.npz'))
    geom, feats = vp.reference_test_inputs()
    kept, lin, pos = vp.cell_index_ref(geom.int(), (128, 128, 1))
    assert int(kept.sum()) == int(fx['kept_count'])
    assert hashlib.sha256(kept.numpy().tobytes()).digest() == fx['kept_sha256'].tobytes()
    assert abs(float(kept.float().mean()) - 0.254) < 2e-3
    out = vp.voxel_pooling_ref(geom.int(), feats, (128, 128, 1)).contiguous()
    rows = out.permute(0, 2, 3, 1).reshape(-1, 80)[torch.from_numpy(fx['probe_cells'])]
    assert np.array_equal(rows.numpy(), fx['probe_rows'])
    assert np.allclose(out.double().sum(dim=(0, 2, 3)).numpy(), fx['channel_sums'], rtol=1e-9)
    assert int((out.abs().sum(1) > 0).sum()) == int(fx['occupied_cells'])
    # pos_memo: (b, y, x) for kept points, -1 otherwise (voxel_pooling_forward_cuda.cu:27-29)
    g = geom.int()
    assert torch.equal(pos[kept][:, 1], g[kept][:, 1]) and torch.equal(pos[kept][:, 2], g[kept][:, 0])
    assert bool((pos[~kept] == -1).all())


def test_truncation_toward_zero_and_z_gate():
    # Appendix B: .int() truncates toward zero; z only gates
    assert torch.tensor([-0.5, -0.99, -1.0, 0.99]).int().tolist() == [0, 0, -1, 0]
    geom = torch.tensor([[[0, 0, 0], [0, 0, 1], [3, 1, 0], [4, 0, 0], [-1, 0, 0], [3, 1, 0]]], dtype=torch.int32)
    feats = torch.arange(6 * 2, dtype=torch.float32).view(1, 6, 2)
    out = vp.voxel_pooling_ref(geom, feats, (4, 2, 1))
    assert out.shape == (1, 2, 2, 4)
    exp = torch.zeros(1, 2, 4, 2)
    exp[0, 0, 0] = feats[0, 0]
    exp[0, 1, 3] = feats[0, 2] + feats[0, 5]
    assert torch.equal(out, exp.permute(0, 3, 1, 2))


def test_backward_oracle_equals_autograd():
    torch.manual_seed(3)
    geom = torch.randint(-2, 10, (2, 300, 3), dtype=torch.int32)
    feats = torch.randn(2, 300, 8, requires_grad=True)
    X, Y, Z = 8, 6, 4
    kept, lin, _ = vp.cell_index_ref(geom, (X, Y, Z))
    k = kept.reshape(-1)
    out = torch.zeros(2 * Y * X, 8).index_add(0, lin.reshape(-1)[k], feats.reshape(-1, 8)[k])
    out = out.view(2, Y, X, 8).permute(0, 3, 1, 2)
    go = torch.randn(2, 8, Y, X)
    out.backward(go)
    ref = vp.voxel_pooling_backward_ref(geom, go, (X, Y, Z), feats.shape)
    assert torch.equal(ref, feats.grad)
    assert torch.equal(out.detach(), vp.voxel_pooling_ref(geom, feats.detach(), (X, Y, Z)))


def test_fused_oracle_forward_and_grads():
    torch.manual_seed(4)
    B, N, D, H, W, C = 2, 2, 5, 3, 4, 8
    geom = torch.randint(-1, 7, (B, N, D, H, W, 3), dtype=torch.int32)
    depth = torch.rand(B * N, D, H, W).softmax(1)
    ctx = torch.rand(B * N, C, H, W) - 0.5
    vn = (6, 5, 3)
    out = vp.voxel_pooling_fused_ref(geom, depth, ctx, vn)
    # brute force
    exp = torch.zeros(B, 5, 6, C)
    for b in range(B):
        for n in range(N):
            for d in range(D):
                for h in range(H):
                    for w in range(W):
                        x, y, z = geom[b, n, d, h, w].tolist()
                        if 0 <= x < 6 and 0 <= y < 5 and 0 <= z < 3:
                            exp[b, y, x] += depth[b * N + n, d, h, w] * ctx[b * N + n, :, h, w]
    assert torch.allclose(out, exp.permute(0, 3, 1, 2), rtol=1e-5, atol=1e-6)
    go = torch.rand(B, C, 5, 6)
    gd, gc = vp.voxel_pooling_fused_grads_ref(geom, depth, ctx, vn, go)
    gd2 = torch.zeros_like(depth, dtype=torch.float64)
    gc2 = torch.zeros_like(ctx, dtype=torch.float64)
    for b in range(B):
        for n in range(N):
            for d in range(D):
                for h in range(H):
                    for w in range(W):
                        x, y, z = geom[b, n, d, h, w].tolist()
                        if 0 <= x < 6 and 0 <= y < 5 and 0 <= z < 3:
                            g = go[b, :, y, x].double()
                            gd2[b * N + n, d, h, w] = (g * ctx[b * N + n, :, h, w].double()).sum()
                            gc2[b * N + n, :, h, w] += depth[b * N + n, d, h, w].double() * g
    assert torch.allclose(gd, gd2, rtol=1e-12, atol=1e-14)
    assert torch.allclose(gc, gc2, rtol=1e-12, atol=1e-14)


def test_run_plan_oracle_is_consistent_with_the_point_oracle():
    """``run_plan_ref`` (the expected contents of a run plan) regroups the reference's terms without
    changing them: summing depth*context run by run and then per cell equals the point-by-point oracle,
    every kept point belongs to exactly one run, and runs never cross a 16-row block."""
    g = torch.Generator().manual_seed(5)
    B, N, D, H, W, C, vn = 2, 2, 6, 20, 5, 8, (16, 8, 1)
    x = torch.randint(-1, 17, (B, N, D, 1, W), generator=g).expand(B, N, D, H, W)
    y = torch.randint(-1, 9, (B, N, D, 1, W), generator=g).expand(B, N, D, H, W)
    z = torch.randint(-1, 2, (B, N, D, H, W), generator=g)
    geom = torch.stack([x, y, z], -1).int().contiguous()
    depth = torch.rand(B * N, D, H, W, generator=g).softmax(1).double()
    ctx = (torch.rand(B * N, C, H, W, generator=g) - 0.5).double()
    head, code, cell_start, sorted_ids = vp.run_plan_ref(geom, vn)
    kept, lin, _ = vp.cell_index_ref(geom, vn)
    assert int((code >= 0).sum()) == sorted_ids.numel() == int(cell_start[-1])
    assert torch.equal((code != -1), kept)
    hh = torch.arange(H).view(1, 1, 1, H, 1).expand_as(head)
    assert bool(head[(hh % 16 == 0) & kept.view_as(head)].all())          # a block's first kept row starts a run
    # run rows, then per-cell sums in slot order
    feats = vp.materialise_features_ref(depth, ctx, B, N).reshape(-1, C)
    code_f = code.reshape(-1)
    run_rows = torch.zeros(sorted_ids.numel(), C, dtype=torch.float64)
    slot = torch.full((code_f.numel(),), -1, dtype=torch.long)
    cur = -1
    Wn = W
    # walk every (b, n, d, w) column top to bottom: a head opens a run, -2 continues it
    idx = torch.arange(code_f.numel()).view(B, N, D, H, W)
    for col in idx.permute(0, 1, 2, 4, 3).reshape(-1, H).tolist():
        cur = -1
        for p in col:
            cv = int(code_f[p])
            if cv >= 0:
                cur = cv
            elif cv == -1:
                cur = -1
            if cv != -1:
                assert cur >= 0
                slot[p] = cur
    k = slot >= 0
    run_rows.index_add_(0, slot[k], feats[k])
    X, Y, _ = vn
    out = torch.zeros(B * X * Y, C, dtype=torch.float64)
    cells = lin.reshape(-1)[sorted_ids]
    out.index_add_(0, cells, run_rows)
    ref = vp.voxel_pooling_ref(geom, feats.view(B, -1, C), vn, acc_dtype=torch.float64)
    assert torch.allclose(out.view(B, Y, X, C).permute(0, 3, 1, 2), ref, rtol=1e-12, atol=1e-14)
