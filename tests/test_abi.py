"""CPU checks of the C-ABI boundary: the library builds, loads, exports every symbol that
``include/bevpool_sm100.h`` declares, and argument errors come back as codes (no compute)."""
import ctypes
import os
import re

import pytest

from mm_training_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def L():
    build.build()
    return _lib.lib()


def _declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'bevpool_sm100.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b((?:bevpool|bevvox|bevlabel|bevdepth|pillar)_[a-z0-9_]+)\s*\(', text)))


def test_every_declared_symbol_is_exported(L):
    names = _declared_symbols()
    assert 'bevpool_forward' in names and 'bevpool_fused_backward' in names
    for n in names:
        assert hasattr(L, n), f'{n} declared in include/bevpool_sm100.h but not exported'
    for n in _lib.EXPORTED_SYMBOLS:
        assert n in names, f'{n} bound in _lib.py but not declared in the header'


def test_reference_launcher_symbol_is_exported(L):
    """ops/voxel_pooling/src/voxel_pooling_forward.cpp:21-22 links against this C++ symbol (Itanium mangling of
    voxel_pooling_forward_kernel_launcher(int x6, const int*, const float*, float*, int*, cudaStream_t))."""
    assert hasattr(L, '_Z37voxel_pooling_forward_kernel_launcheriiiiiiPKiPKfPfPiP11CUstream_st')


def test_abi_version_and_error_strings(L):
    assert L.bevpool_abi_version() == 2
    assert L.bevpool_error_string(0) == b'ok'
    assert b'channel' in L.bevpool_error_string(-3)


def test_argument_errors_are_codes_not_crashes(L):
    pb, tb = ctypes.c_size_t(), ctypes.c_size_t()
    assert L.bevpool_plan_sizes(0, 10, 4, 4, ctypes.byref(pb), ctypes.byref(tb)) == -1
    assert L.bevpool_plan_sizes(4, 2 ** 30, 4, 4, ctypes.byref(pb), ctypes.byref(tb)) == -2
    assert L.bevpool_plan_sizes(2, 6000, 128, 128, ctypes.byref(pb), ctypes.byref(tb)) == 0
    assert pb.value > 2 * 6000 * 8 and tb.value > 0
    # null pointers / bad channel counts are rejected before any launch
    assert L.bevpool_forward(None, None, None, 0, 2, 6000, 80, 128, 128, None, None) == -1
    assert L.bevpool_forward(None, None, None, 0, 2, 6000, 81, 128, 128, None, None) == -3
    assert L.bevpool_transpose(None, None, 0, 1, 4, 4, None) == -1


def test_python_ops_refuse_cpu_tensors():
    import torch
    from mm_training_b200.ops.voxel_pooling import voxel_pooling
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        voxel_pooling(torch.zeros(1, 4, 3, dtype=torch.int32), torch.zeros(1, 4, 8), [2, 2, 1])


def test_bench_byte_model_matches_the_survey():
    """bench.py's roofline numerator is SURVEY.md section 8(d), row (B): 32.5 MB per CFG-2 frame."""
    import importlib.util
    spec = importlib.util.spec_from_file_location('bench_mod', os.path.join(ROOT, 'bench.py'))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    from mm_training_b200.configs import CFG_2, CFG_AIM
    b = bench.algorithmic_bytes(CFG_2, 150442)
    assert abs(b['step'] - 32.5e6) < 0.1e6
    assert b['fused_forward'] + b['fused_backward'] < b['step'] * 1.05
    assert abs(bench.algorithmic_bytes(CFG_AIM, 738884)['step'] - 108e6) < 1.5e6
