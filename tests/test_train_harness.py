"""CPU checks of the config-4 harness (bench_train.py): the stand-in network's tensor contract around the hot path
(depth (B*N, D, h, w) fp32 probabilities, context (B*N, C, h, w), pooled grid (B, C, Y, X)) and the restated depth
loss (mm_training_aim.py:163-176)."""
import torch
import torch.nn.functional as F

import bench_train as bt
from mm_training_b200.configs import CFG_2, CFG_3


def test_standin_network_contract_around_the_hot_path():
    torch.manual_seed(0)
    net = bt.StandInFusionNet(CFG_2, CFG_3)
    X, Y, _ = CFG_2.voxel_num
    h, w = CFG_2.feat_hw
    seen = {}

    def pool(depth, context):
        seen['depth'], seen['context'] = depth, context
        return depth.sum() * 0 + context.mean() + torch.zeros(1, CFG_2.output_channels, Y, X)
    preds, depth = net(torch.randn(1, 4, 3, *CFG_2.final_dim), pool, torch.randn(1, CFG_3.vfe_features, 256, 2048))
    assert seen['depth'].shape == (4, CFG_2.depth_bins, h, w) and seen['context'].shape == (4, CFG_2.output_channels, h, w)
    assert torch.allclose(depth.sum(1), torch.ones(4, h, w), atol=1e-5)            # softmax over D (lss_fpn.py:423)
    assert preds.shape == (1, net.num_classes + 8, Y, X)
    preds.sum().backward()
    assert net.depth_net[-1].weight.grad is not None


def test_depth_loss_restatement():
    D = 7
    g = torch.Generator().manual_seed(1)
    preds = torch.rand(2, D, 3, 5, generator=g).softmax(1)
    labels = F.one_hot(torch.randint(0, D, (2 * 3 * 5,), generator=g), D).float()
    labels[::4] = 0                                                                 # pixels without a LiDAR return
    flat = preds.permute(0, 2, 3, 1).reshape(-1, D)
    fg = labels.max(1).values > 0
    want = 3.0 * F.binary_cross_entropy(flat[fg], labels[fg], reduction='none').sum() / max(1.0, float(fg.sum()))
    assert torch.allclose(bt.depth_loss_ref(labels, preds, D), want)
