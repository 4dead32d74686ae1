import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
from test_gpu_nchw_rig import _case
from mm_training_b200.ops.voxel_pooling import voxel_pooling_fused
from oracle import voxel_pool_ref as vp

def run(shape, coherent):
    B, N, D, H, W, C, vn = shape
    geom, depth, ctx, go = _case(21, B, N, D, H, W, C, vn, coherent)
    d = depth.cuda().requires_grad_(True); c = ctx.cuda().requires_grad_(True)
    out = voxel_pooling_fused(geom.cuda(), d, c, vn)
    out.backward(go.cuda())
    gd, gc = vp.voxel_pooling_fused_grads_ref(geom, depth, ctx, vn, go)
    e = (d.grad.double().cpu() - gd).abs()
    bad = (e > 1e-4).nonzero()
    e2 = (c.grad.double().cpu() - gc).abs()
    print(shape, coherent, 'gd bad', bad.shape[0], 'of', e.numel(), 'gc bad', int((e2 > 1e-3).sum()))
    if bad.shape[0]:
        print('  first bad (bn,d,h,w):', bad[:12].tolist())
        import collections
        print('  by h:', sorted(collections.Counter(bad[:, 2].tolist()).items()))
        print('  by d:', sorted(collections.Counter(bad[:, 1].tolist()).items())[:40])
        print('  by w:', sorted(collections.Counter(bad[:, 3].tolist()).items()))
        k = bad[0].tolist()
        print('  got', float(d.grad[tuple(k)]), 'want', float(gd[tuple(k)]))

for shape in [(1, 3, 37, 44, 12, 32, (40, 12, 1)), (1, 1, 37, 16, 12, 32, (40, 12, 1)), (1, 1, 32, 44, 12, 32, (40, 12, 1)),
              (1, 1, 37, 44, 12, 80, (40, 12, 1)), (1, 1, 16, 32, 4, 32, (40, 12, 1))]:
    for coh in (True, False):
        run(shape, coh)
