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
