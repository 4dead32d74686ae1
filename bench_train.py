#!/usr/bin/env python
"""Config-4 harness (BASELINE.json configs[3]): a LiDAR + radar + camera fusion TRAINING STEP with the fused
pooling op and the native voxelizer inside a real autograd graph, bf16 autocast, NCCL DDP -- one process per GPU.

    python bench.py --workload train [--gpus N] [--steps K] [--warmup W] [--train-cfg aim|cfg2] [--batch B]

The reference's step is ``exps/mm_training_aim.py:252-289`` -> ``models/bev_depth.py:163-200`` (BEVDepthLiDAR).  Its
dense parts (ResNet-50 + FPN + DepthNet, mmdet3d's SparseEncoder, the CenterPoint head) are third-party models that
are neither vendored in the reference nor installed here, and they are OUT OF SCOPE of this repository (DESIGN.md
section 7): they stay stock PyTorch / cuDNN.  This harness therefore uses a dependency-free STAND-IN network of the
same topology and tensor shapes around the hot path:

    images (B, N, 3, H, W) -> torchvision ResNet-50 (random init, to stride 16) -> 1x1 neck -> DepthNet stand-in
        -> depth logits (B*N, D, h, w) + context (B*N, C, h, w) -> softmax (lss_fpn.py:423)
        -> [HOT PATH] camera pooling: geometry + index + outer product + voxel pooling (lss_fpn.py:441-466)
    clouds  list of (Np, 8) LiDAR+radar points
        -> [HOT PATH] hard voxelization -> HardSimpleVFE mean -> dense scatter (bev_depth.py:181-183)
        -> two strided convs standing in for SparseEncoder's conv stack -> (B, 256, Y, X)
    cat -> BEVFuseLayer (restated from bev_depth.py:133-145) -> conv head -> heat-map / box losses
    + the depth loss of mm_training_aim.py:163-176 on the returned depth probabilities.

Two arms in the SAME harness, same weights, same inputs:
    fused      this repository's ops: run plan from the camera rig + voxel_pooling_fused (fp32 at the op boundary,
               exactly like the reference, whose kernel only takes float32: softmax runs in fp32 under autocast and
               fp32 x bf16 promotes to fp32)
    reference  the reference's pipeline: torch geometry ops + materialised outer product + permute + contiguous +
               its own CUDA kernel (oracle/_ref, compiled unmodified) + its autograd backward
The voxelizer is the native one in both arms (mmcv is not installed; its CPU kernel is timed in bench.py's lidar block).
"""
from __future__ import annotations

import json
import os
import sys
import time

import torch
import torch.distributed as dist
import torch.nn.functional as F
from torch import nn

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


class BEVFuseLayer(nn.Module):
    """Restated from models/bev_depth.py:133-145: 3x3 conv, then a squeeze-excite style channel gate."""

    def __init__(self, in_channels):
        super().__init__()
        self.conv_3 = nn.Conv2d(in_channels, in_channels, 3, padding=1)
        self.conv_1 = nn.Conv2d(in_channels, in_channels, 1)
        self.avg_pool = nn.AdaptiveAvgPool2d((1, 1))

    def forward(self, x):
        x = self.conv_3(x)
        return x * torch.sigmoid(self.conv_1(self.avg_pool(x)))


class StandInFusionNet(nn.Module):
    def __init__(self, cam_cfg, vox_cfg, lidar_channels=256, num_classes=10):
        super().__init__()
        import torchvision
        r = torchvision.models.resnet50(weights=None)
        self.img_backbone = nn.Sequential(r.conv1, r.bn1, r.relu, r.maxpool, r.layer1, r.layer2, r.layer3)   # stride 16, 1024 ch
        self.neck = nn.Sequential(nn.Conv2d(1024, 512, 1, bias=False), nn.BatchNorm2d(512), nn.ReLU(True))
        self.D, self.C = cam_cfg.depth_bins, cam_cfg.output_channels
        self.depth_net = nn.Sequential(nn.Conv2d(512, 256, 3, padding=1, bias=False), nn.BatchNorm2d(256), nn.ReLU(True),
                                       nn.Conv2d(256, self.D + self.C, 1))
        self.vfe_features = vox_cfg.vfe_features
        self.lidar_encoder = nn.Sequential(nn.Conv2d(self.vfe_features, 64, 3, stride=2, padding=1), nn.ReLU(True),
                                           nn.Conv2d(64, lidar_channels, 3, stride=2, padding=1), nn.ReLU(True))
        self.bev_fuse = BEVFuseLayer(self.C + lidar_channels)
        self.head = nn.Sequential(nn.Conv2d(self.C + lidar_channels, 64, 3, padding=1), nn.ReLU(True),
                                  nn.Conv2d(64, num_classes + 8, 1))
        self.num_classes = num_classes

    def forward(self, imgs, pool_fn, lidar_canvas):
        B, N = imgs.shape[:2]
        x = self.neck(self.img_backbone(imgs.flatten(0, 1).contiguous(memory_format=torch.channels_last)))
        depth_feature = self.depth_net(x)
        depth = depth_feature[:, :self.D].softmax(1)                      # lss_fpn.py:423 (fp32 under autocast)
        context = depth_feature[:, self.D:self.D + self.C]
        img_bev = pool_fn(depth, context)                                 # (B, C, Y, X)
        lidar_bev = self.lidar_encoder(lidar_canvas)
        if lidar_bev.shape[-2:] != img_bev.shape[-2:]:
            lidar_bev = F.interpolate(lidar_bev, size=img_bev.shape[-2:])
        bev = self.bev_fuse(torch.cat([img_bev.to(lidar_bev.dtype), lidar_bev], 1))
        return self.head(bev), depth


def depth_loss_ref(depth_labels, depth_preds, D):
    """mm_training_aim.py:163-176 (get_depth_loss), restated."""
    depth_labels = depth_labels.view(-1, D)
    depth_preds = depth_preds.permute(0, 2, 3, 1).contiguous().view(-1, D)
    fg_mask = torch.max(depth_labels, dim=1).values > 0.0
    with torch.autocast('cuda', enabled=False):
        loss = F.binary_cross_entropy(depth_preds[fg_mask].float(), depth_labels[fg_mask], reduction='none').sum() / \
            max(1.0, fg_mask.sum())
    return 3.0 * loss


def run_train(args, rank, world, local_rank, json_fd):
    from mm_training_b200 import synthetic
    from mm_training_b200.configs import CFG_2, CFG_3, CFG_AIM
    from mm_training_b200.ops.voxel_pooling import LiftSplatGeometry, rig_variant, voxel_pooling_fused
    from mm_training_b200.ops.voxelize import Voxelization, voxelize
    from mm_training_b200 import _lib
    from bench import ClockSampler
    cfg = CFG_AIM if args.train_cfg == 'aim' else CFG_2
    vox = CFG_3
    B = args.batch if args.batch_given else (1 if cfg is CFG_AIM else 2)
    dev = torch.device('cuda', local_rank)
    torch.manual_seed(7)                                    # same initial weights on every rank and in both arms
    net = StandInFusionNet(cfg, vox).to(dev).to(memory_format=torch.channels_last)
    state0 = {k: v.clone() for k, v in net.state_dict().items()}
    model = nn.parallel.DistributedDataParallel(net, device_ids=[local_rank]) if world > 1 else net
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4)

    g = torch.Generator().manual_seed(11 + rank)
    H_img, W_img = cfg.final_dim
    imgs = torch.randn(B, cfg.num_cams, 3, H_img, W_img, generator=g).to(dev)
    s2e, intrin = synthetic.camera_rig_mats(cfg, B, device=dev, yaw_jitter_deg=5.0, seed=11 + rank)
    clouds = [torch.from_numpy(synthetic.lidar_sweep(vox.points_per_sweep, 8, seed=100 * rank + i)).to(dev) for i in range(B)]
    layer = Voxelization(list(vox.voxel_size), list(vox.point_cloud_range), vox.max_num_points, vox.max_voxels)
    h, w = cfg.feat_hw
    lab_idx = torch.randint(0, cfg.depth_bins, (B * cfg.num_cams * h * w,), generator=g)
    depth_labels = F.one_hot(lab_idx, cfg.depth_bins).float().to(dev)
    X, Y, _ = cfg.voxel_num
    heat_t = torch.rand(B, net.num_classes, Y, X, generator=g).pow(8).to(dev)
    lsg = LiftSplatGeometry.from_config(cfg, dev)
    variant = rig_variant(dev)
    vn = lsg.voxel_num

    hint = {'n': None, 'plans': []}

    def pool_fused(depth, context):
        # the rig of this rank is fixed: the run count is read back once, later plans are sync-free (status checked below)
        plan = lsg.plan(s2e, intrin, max_runs=hint['n'], variant=variant)
        if hint['n'] is None and plan.mode == 'runs':
            hint['n'] = int(plan.num_sorted * 1.125) + 1024
        hint['plans'] = [plan]
        return voxel_pooling_fused(None, depth.float().contiguous(), context.float().contiguous(), vn, plan)

    def pool_reference(depth, context):
        from oracle import ref_cuda_op
        geom = lsg.geom_xyz(s2e, intrin)                     # lss_fpn.py:455-462 in torch, like the reference
        return ref_cuda_op.ref_pipeline(geom, depth, context, vn)     # fp32 x bf16 -> fp32, like the reference

    def make_step(pool_fn):
        def step():
            with torch.autocast('cuda', dtype=torch.bfloat16):
                canvas = voxelize(clouds, layer, mean_features=vox.vfe_features, padded=True, scatter=True)[4]
                preds, depth = model(imgs, pool_fn, canvas)
                heat, reg = preds[:, :net.num_classes], preds[:, net.num_classes:]
                loss = F.binary_cross_entropy_with_logits(heat.float(), heat_t) + reg.float().abs().mean() * 0.25
                loss = loss + depth_loss_ref(depth_labels, depth, cfg.depth_bins)
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
            return loss.detach()
        return step

    def timed(pool_fn, steps, warmup):
        net.load_state_dict(state0)
        step = make_step(pool_fn)
        for _ in range(max(3, warmup)):
            last = step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        l0 = _lib.launch_count()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local_rank) as clocks:
            a.record()
            for _ in range(steps):
                last = step()
            b.record()
            torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), float(last.item()), _lib.launch_count() - l0, clocks.summary()

    fused_ms, fused_loss, launches, clocks = timed(pool_fused, args.steps, args.warmup)
    assert hint['plans'][0].status() == 0, 'run-row scratch overflow'

    # ---- end to end: the step's inputs start in pinned host memory every step, the loss is read back
    h_imgs = imgs.cpu().pin_memory()
    h_clouds = [c.cpu().pin_memory() for c in clouds]
    h_s2e, h_k = s2e.cpu().pin_memory(), intrin.cpu().pin_memory()
    e2e_step_fn = make_step(pool_fused)

    def e2e_step():
        imgs.copy_(h_imgs, non_blocking=True)
        for c, hc in zip(clouds, h_clouds):
            c.copy_(hc, non_blocking=True)
        s2e.copy_(h_s2e, non_blocking=True)
        intrin.copy_(h_k, non_blocking=True)
        return float(e2e_step_fn().item())
    for _ in range(2):
        e2e_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e2e_steps = max(3, min(args.steps, 10))
    ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ea.record()
    for _ in range(e2e_steps):
        e2e_step()
    eb.record()
    torch.cuda.synchronize()
    te = torch.tensor([ea.elapsed_time(eb)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_ms = float(te.item()) / e2e_steps
    h2d = sum(t.numel() * t.element_size() for t in [h_imgs, h_s2e, h_k] + h_clouds)
    ref = None
    try:
        from oracle import ref_cuda_op
        if ref_cuda_op.available():
            ref_steps = max(3, min(args.steps, 10))
            ref_ms, ref_loss, _, _ = timed(pool_reference, ref_steps, 3)
            ref = {'what': "same harness, the reference's camera pooling pipeline (torch geometry + materialised outer product + "
                           "its own CUDA kernel from oracle/_ref + its autograd backward)",
                   'iter_per_s': ref_steps / (ref_ms * 1e-3), 'ms_per_step': ref_ms / ref_steps, 'loss_after': ref_loss}
    except Exception as e:                                      # pragma: no cover
        ref = {'error': repr(e)}
    if rank != 0:
        return
    ms = fused_ms / args.steps
    line = {'metric': 'fusion_train_step_iter_per_s', 'value': 1e3 / ms, 'unit': 'iter/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': max(3, args.warmup), 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'bf16 autocast (fp32 at the pooling op and the voxelizer, like the reference)', 'data': 'synthetic',
            'config': {'workload': 'train_standin_fusion_' + cfg.name, 'frames_per_gpu_per_step': B, 'global_batch': B * world,
                       'parallelism': f'ddp{world} (NCCL gradient all-reduce; the ops shard by sample, no collective)',
                       'network': 'stand-in: torchvision ResNet-50 (to stride 16) + 1x1 neck + DepthNet stand-in, native voxelizer '
                                  '-> VFE mean -> scatter -> 2 strided convs, BEVFuseLayer, conv head; AdamW',
                       'lidar': f'{vox.points_per_sweep} LiDAR+radar points x 8 features per frame, {vox.name}',
                       'geometry': 'rig (on-device, variant %s)' % variant if variant is not None else 'torch geometry ops'},
            'frames_per_s': B * world * 1e3 / ms, 'loss_after': fused_loss, 'gpu_launches': launches, 'clocks': clocks,
            'reference_pipeline_same_harness': ref,
            'speedup_vs_reference_pipeline': (ref['ms_per_step'] / ms) if ref and 'ms_per_step' in ref else None,
            'e2e': {'value': 1e3 / e2e_ms, 'unit': 'iter/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 4,
                    'ms_per_step': e2e_ms, 'steps': e2e_steps,
                    'api': 'images, clouds and rig matrices copied from pinned host memory every step; loss.item() read back'}}
    os.write(json_fd, (json.dumps(line) + '\n').encode())
