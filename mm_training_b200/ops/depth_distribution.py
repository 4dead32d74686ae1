"""Depth distribution of the camera branch: softmax over D + depth-oracle overwrite in one native pass
(csrc/depth_softmax.cu), differentiable.

Replaces ``layers/backbones/lss_fpn.py:423`` and ``:427-434`` of the reference::

    depth = depth_feature[:, :D].softmax(1)
    if depth_oracle is not None:            # pixels with a LiDAR return take the oracle's distribution
        fg_mask = (torch.max(depth_oracle, dim=1).values > 0.0).view(-1)
        ... permute / contiguous / masked index_put / permute ...

``depth_distribution(depth_feature, depth_channels, depth_oracle=None) -> (depth, depth_used)``: ``depth`` is the
softmax (returned for the depth loss, ``lss_fpn.py:466``), ``depth_used`` what the pooling consumes (the same tensor
when there is no oracle).  Both are float32 (B*N, D, H, W) contiguous -- the layout the fused pooling kernels read --
whatever the dtype of ``depth_feature`` (autocast runs softmax in float32 too).  No CPU fallback.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch.autograd import Function

from .. import _lib


def _forward(logits: torch.Tensor, oracle: Optional[torch.Tensor]):
    _lib.require_cuda(logits)
    BN, D, H, W = logits.shape
    if logits.stride()[1:] != (H * W, W, 1):            # a channel slice of an NCHW tensor passes; anything else is copied
        logits = logits.contiguous()
    dev = logits.device
    with torch.cuda.device(dev):
        prob = torch.empty(BN, D, H, W, dtype=torch.float32, device=dev)
        used = None
        if oracle is not None:
            _lib.require_cuda(oracle)
            assert oracle.shape == (BN, D, H, W)
            oracle = oracle.float().contiguous()
            used = torch.empty_like(prob)
        _lib.check(_lib.lib().bevdepth_softmax_forward(
            logits.data_ptr(), _lib.dtype_code(logits), logits.stride(0), oracle.data_ptr() if oracle is not None else None,
            BN, D, H, W, prob.data_ptr(), used.data_ptr() if used is not None else None, _lib.stream_ptr(dev)),
            'bevdepth_softmax_forward')
    return prob, used, oracle


def _backward(prob, oracle, grad_prob, grad_used, dtype):
    BN, D, H, W = prob.shape
    dev = prob.device
    gp = grad_prob.float().contiguous() if grad_prob is not None else None
    gu = grad_used.float().contiguous() if grad_used is not None else None
    with torch.cuda.device(dev):
        grad_logits = torch.empty(BN, D, H, W, dtype=dtype, device=dev)
        if gp is None and gu is None:
            return grad_logits.zero_()
        _lib.check(_lib.lib().bevdepth_softmax_backward(
            prob.data_ptr(), gp.data_ptr() if gp is not None else None, gu.data_ptr() if gu is not None else None,
            oracle.data_ptr() if oracle is not None else None, BN, D, H, W, grad_logits.data_ptr(),
            _lib.dtype_code(grad_logits), _lib.stream_ptr(dev)), 'bevdepth_softmax_backward')
    return grad_logits


class _DepthSoftmax(Function):                               # no oracle: one output, consumed by the loss AND the pooling
    @staticmethod
    def forward(ctx, logits):
        prob, _, _ = _forward(logits, None)
        ctx.save_for_backward(prob)
        ctx.in_dtype = logits.dtype
        return prob

    @staticmethod
    def backward(ctx, grad_prob):
        (prob,) = ctx.saved_tensors
        return _backward(prob, None, grad_prob, None, ctx.in_dtype)


class _DepthSoftmaxOracle(Function):
    @staticmethod
    def forward(ctx, logits, oracle):
        prob, used, oracle = _forward(logits, oracle)
        ctx.save_for_backward(prob, oracle)
        ctx.in_dtype = logits.dtype
        return prob, used

    @staticmethod
    def backward(ctx, grad_prob, grad_used):
        prob, oracle = ctx.saved_tensors
        return _backward(prob, oracle, grad_prob, grad_used, ctx.in_dtype), None


def depth_distribution(depth_feature: torch.Tensor, depth_channels: int,
                       depth_oracle: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """depth_feature (B*N, >= depth_channels, H, W) -> (depth, depth_used), both float32 (B*N, depth_channels, H, W);
    without an oracle they are the same tensor."""
    logits = depth_feature[:, :depth_channels]
    if depth_oracle is None:
        p = _DepthSoftmax.apply(logits)
        return p, p
    return _DepthSoftmaxOracle.apply(logits, depth_oracle)
