"""GPU tests of the drop-in boundary itself: the reference's native entry point exported by libbevpool_sm100
(``voxel_pooling_forward_kernel_launcher``, ops/voxel_pooling/src/voxel_pooling_forward_cuda.cu:38-42) and the
ctypes stub of INTEGRATION.md section 2(b), executed verbatim, both on the reference's own unit-test recipe
(test/test_ops/test_voxel_pooling.py:15-37)."""
import ctypes
import os
import re
import types

import pytest
import torch

from mm_training_b200 import _lib
from oracle import ref_cuda_op
from oracle import voxel_pool_ref as vp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _reference_forward(ext):
    """ops/voxel_pooling/voxel_pooling.py:30-55 of the reference around ``ext.voxel_pooling_forward_wrapper``."""
    geom_xyz, features = vp.reference_test_inputs()
    g, f = geom_xyz.cuda().int().contiguous(), features.cuda().contiguous()
    B, C = f.shape[0], f.shape[-1]
    g = g.reshape(B, -1, 3)
    f = f.reshape(B, -1, C)
    out = f.new_zeros(B, 128, 128, C)
    pos_memo = g.new_ones(B, g.shape[1], 3) * -1
    assert ext.voxel_pooling_forward_wrapper(B, g.shape[1], C, 128, 128, 1, g, f, out, pos_memo) == 1
    gold = vp.python_loop_golden(geom_xyz, features, (128, 128, 1))
    assert torch.allclose(gold.cuda(), out.permute(0, 3, 1, 2), rtol=1e-3)          # the reference test's own bar
    _, _, pos = vp.cell_index_ref(geom_xyz.int(), (128, 128, 1))
    assert torch.equal(pos_memo.cpu(), pos.reshape(B, -1, 3))                        # integer indices: bit-exact
    return out, pos_memo


def test_integration_md_stub_runs_the_reference_recipe():
    text = open(os.path.join(ROOT, 'INTEGRATION.md')).read()
    block = re.search(r"```python\n(# ops/voxel_pooling/voxel_pooling_ext.py.*?)```", text, flags=re.S).group(1)
    block = block.replace("'libbevpool_sm100.so'", repr(_lib.LIB_PATH))
    ext = types.ModuleType('voxel_pooling_ext')
    exec(compile(block, 'INTEGRATION.md', 'exec'), ext.__dict__)
    out, pos = _reference_forward(ext)
    if ref_cuda_op.available():                                                      # and against the reference's own kernel
        geom_xyz, features = vp.reference_test_inputs()
        ref_out, ref_pos = ref_cuda_op.ref_forward_with_pos_memo(geom_xyz.cuda().int(), features.cuda(), [128, 128, 1])
        assert torch.equal(pos.view_as(ref_pos), ref_pos)
        assert torch.allclose(ref_out, out.permute(0, 3, 1, 2), rtol=1e-5, atol=2e-6)


def test_reference_launcher_symbol():
    # the C++ symbol the reference's voxel_pooling_forward.cpp:21-22 links against, called like :34 does
    L = ctypes.CDLL(_lib.LIB_PATH)
    fn = getattr(L, '_Z37voxel_pooling_forward_kernel_launcheriiiiiiPKiPKfPfPiP11CUstream_st')
    fn.restype = None
    fn.argtypes = [ctypes.c_int] * 6 + [ctypes.c_void_p] * 5

    class Ext:
        @staticmethod
        def voxel_pooling_forward_wrapper(b, n, c, x, y, z, geom, feats, out, pos):
            fn(b, n, c, x, y, z, geom.data_ptr(), feats.data_ptr(), out.data_ptr(), pos.data_ptr(),
               torch.cuda.current_stream().cuda_stream)
            return 1
    a, pa = _reference_forward(Ext)
    b, pb = _reference_forward(Ext)
    assert torch.equal(a, b) and torch.equal(pa, pb)                                 # bit-stable run to run
