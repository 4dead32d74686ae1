"""The reference's own CUDA voxel_pooling kernel, run on the GPU box as an on-device oracle and
as the "reference CUDA op" timing arm.  TEST / BENCH INFRASTRUCTURE -- never imported by the product.

``oracle/Makefile`` (target ``ref``) compiles the reference's translation unit
``ops/voxel_pooling/src/voxel_pooling_forward_cuda.cu`` for sm_100a, unmodified and from where it
lies in the reference checkout, into ``oracle/_ref/libref_voxel_pooling.so``.  Its launcher is a
plain C++ symbol (``voxel_pooling_forward_kernel_launcher``, :38-42), called here through ctypes
by its mangled name.  The python around it restates the reference's autograd wrapper
(``ops/voxel_pooling/voxel_pooling.py:10-69``): zero output + ``pos_memo = -1`` allocation, the
permuted return view, and the mask / advanced-index backward.
"""
from __future__ import annotations

import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB = os.path.join(_HERE, '_ref', 'libref_voxel_pooling.so')
_SYMBOL = '_Z37voxel_pooling_forward_kernel_launcheriiiiiiPKiPKfPfPiP11CUstream_st'
_fn = None


def available() -> bool:
    return os.path.exists(REF_LIB)


def _launcher():
    global _fn
    if _fn is None:
        lib = ctypes.CDLL(REF_LIB)
        fn = getattr(lib, _SYMBOL)
        fn.restype = None
        fn.argtypes = [ctypes.c_int] * 6 + [ctypes.c_void_p] * 5
        _fn = fn
    return _fn


class RefVoxelPooling(torch.autograd.Function):
    @staticmethod
    def forward(ctx, geom_xyz, input_features, voxel_num):
        assert geom_xyz.is_contiguous() and input_features.is_contiguous()
        assert input_features.dtype == torch.float32 and geom_xyz.dtype == torch.int32
        grad_input_features = torch.zeros_like(input_features)                  # voxel_pooling.py:29
        geom_xyz = geom_xyz.reshape(geom_xyz.shape[0], -1, geom_xyz.shape[-1])
        input_features = input_features.reshape(geom_xyz.shape[0], -1, input_features.shape[-1])
        assert geom_xyz.shape[1] == input_features.shape[1]
        batch_size, num_points, num_channels = input_features.shape
        vx, vy, vz = (int(v) for v in voxel_num)
        output_features = input_features.new_zeros(batch_size, vy, vx, num_channels)
        pos_memo = geom_xyz.new_ones(batch_size, num_points, 3) * -1              # voxel_pooling.py:40
        _launcher()(batch_size, num_points, num_channels, vx, vy, vz, geom_xyz.data_ptr(),
                    input_features.data_ptr(), output_features.data_ptr(), pos_memo.data_ptr(),
                    torch.cuda.current_stream().cuda_stream)
        ctx.save_for_backward(grad_input_features, pos_memo)
        ctx.pos_memo = pos_memo
        return output_features.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, grad_output_features):
        grad_input_features, pos_memo = ctx.saved_tensors
        kept = (pos_memo != -1)[..., 0]
        shape = grad_input_features.shape
        grad_input_features = grad_input_features.reshape(shape[0], -1, shape[-1])
        grad_input_features[kept] = grad_output_features[
            pos_memo[kept][..., 0].long(), :, pos_memo[kept][..., 1].long(), pos_memo[kept][..., 2].long()]
        return None, grad_input_features.reshape(shape), None


def ref_voxel_pooling(geom_xyz, input_features, voxel_num):
    if isinstance(voxel_num, torch.Tensor):
        voxel_num = voxel_num.tolist()
    return RefVoxelPooling.apply(geom_xyz, input_features, voxel_num)


def ref_forward_with_pos_memo(geom_xyz, input_features, voxel_num):
    """(out (B,C,Y,X), pos_memo (B,Np,3)) straight from the reference kernel."""
    if isinstance(voxel_num, torch.Tensor):
        voxel_num = voxel_num.tolist()
    g = geom_xyz.reshape(geom_xyz.shape[0], -1, 3)
    f = input_features.reshape(g.shape[0], -1, input_features.shape[-1])
    b, n, c = f.shape
    vx, vy, vz = (int(v) for v in voxel_num)
    out = f.new_zeros(b, vy, vx, c)
    pos = g.new_ones(b, n, 3) * -1
    _launcher()(b, n, c, vx, vy, vz, g.data_ptr(), f.data_ptr(), out.data_ptr(), pos.data_ptr(),
                torch.cuda.current_stream().cuda_stream)
    return out.permute(0, 3, 1, 2), pos


def ref_pipeline(geom_xyz, depth, context, voxel_num):
    """The reference's camera pooling pipeline exactly as ``layers/backbones/lss_fpn.py:441-466``:
    outer product -> reshape -> permute -> ``.contiguous()`` -> op -> ``.contiguous()``."""
    B, N = geom_xyz.shape[0], geom_xyz.shape[1]
    f = depth.unsqueeze(1) * context.unsqueeze(2)
    f = f.reshape(B, N, f.shape[1], f.shape[2], f.shape[3], f.shape[4]).permute(0, 1, 3, 4, 5, 2)
    return ref_voxel_pooling(geom_xyz, f.contiguous(), voxel_num).contiguous()
