import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mm_training_b200 import synthetic
from mm_training_b200.configs import CFG_2
from mm_training_b200.ops.voxel_pooling import build_plan, voxel_pooling, voxel_pooling_fused
from oracle import voxel_pool_ref as vp
cfg, B = CFG_2, 2
geom, vn = synthetic.camera_rig(cfg, B, yaw_jitter_deg=5.0)
depth, ctx, go = synthetic.camera_features(cfg, B)
vn = vn.tolist()
feats = vp.materialise_features_ref(depth, ctx, B, cfg.num_cams)
ref = vp.voxel_pooling_ref(geom, feats, vn)
fused = voxel_pooling_fused(geom.cuda(), depth.cuda(), ctx.cuda(), vn).cpu()
drop = voxel_pooling(geom.cuda(), feats.cuda(), vn).cpu()
feats_gpu = vp.materialise_features_ref(depth.cuda(), ctx.cuda(), B, cfg.num_cams)
print('feats cpu==gpu', torch.equal(feats, feats_gpu.cpu()))
drop2 = voxel_pooling(geom.cuda(), feats_gpu, vn).cpu()
for name, t in [('fused', fused), ('drop', drop), ('drop_gpufeats', drop2)]:
    d = (t - ref).abs()
    print(name, 'equal', torch.equal(t, ref), 'max', float(d.max()), 'n_diff', int((d > 0).sum()))
d = (drop - ref).abs()
if d.max() > 0:
    idx = torch.nonzero(d > 0)
    print(idx[:10], idx.shape)
    b, c, y, x = idx[0].tolist()
    print(drop[b, :, y, x][:8], ref[b, :, y, x][:8])
    plan = build_plan(geom.cuda(), vn)
    cs = plan.cell_start.cpu()
    cell = b * vn[0] * vn[1] + y * vn[0] + x
    print('cell count', int(cs[cell + 1] - cs[cell]))
    ys = torch.unique(idx[:, 2]); xs = torch.unique(idx[:, 3]); print('ys', ys[:20], 'xs', xs[:20], 'cs', torch.unique(idx[:,1])[:20])
