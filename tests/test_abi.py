"""The C-ABI boundary: libdfb200.so loads, exports every symbol include/dfb200.h declares, and the
product path fails loudly (no CPU fallback) when there is no GPU."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
import deepflows_b200


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "dfb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"DFB_API\s+[^;(]*?\b(dfb_\w+)\s*\(", text)))


def test_header_declares_the_protocol():
    syms = _declared_symbols()
    assert len(syms) >= 70
    for name in ("dfb_fill", "dfb_compact", "dfb_ewise_setitem", "dfb_scalar_setitem", "dfb_matmul", "dfb_reduce_sum",
                 "dfb_reduce_max", "dfb_from_host", "dfb_to_host", "dfb_conv2d_fprop", "dfb_conv2d_dgrad",
                 "dfb_conv2d_wgrad", "dfb_bn_fwd_train", "dfb_bn_bwd", "dfb_multi_adam_step", "dfb_multi_sgd_step",
                 "dfb_comm_allreduce_async"):
        assert name in syms


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(deepflows_b200.lib_path())
    missing = [s for s in _declared_symbols() if not hasattr(lib, s)]
    assert not missing, "declared in include/dfb200.h but not exported: %s" % missing
    lib.dfb_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.dfb_version()


def test_shim_imports_at_the_reference_path():
    from DeepFlows.backend.backend_src.build.Release import CUDA_BACKEND as m
    for name in ("Array", "fill", "from_numpy", "to_numpy", "compact", "ewise_setitem", "scalar_setitem", "ewise_add",
                 "ewise_mul", "ewise_div", "ewise_maximum", "ewise_eq", "ewise_ge", "scalar_add", "scalar_mul",
                 "scalar_div", "scalar_power", "scalar_maximum", "scalar_eq", "scalar_ge", "ewise_log", "ewise_exp",
                 "ewise_tanh", "matmul", "reduce_sum", "reduce_max"):
        assert hasattr(m, name), name
    assert m.__max_dimensions__ == 8
    assert m.__device__name__ == "cuda" and isinstance(m.__tile_size__, int) and isinstance(m.__version__, str)
    # the helper containers of the reference module (cu:526-556)
    for cls, bad in ((m.Int32Vector, 2 ** 40), (m.SizeTVector, -1)):
        v = cls()
        for x in (3, 1, 4):
            v.push_back(x)
        assert v.size() == 3 and len(v) == 3 and v[0] == 3 and v[-1] == 4 and list(v) == [3, 1, 4] and tuple(v) == (3, 1, 4)
        with pytest.raises(IndexError):
            v[3]
        with pytest.raises(IndexError):
            v[-4]
        with pytest.raises(TypeError):
            v.push_back(bad)
        v.clear()
        assert v.size() == 0
    shape = m.Int32Vector()
    shape.push_back(2)
    with pytest.raises(ValueError, match="cannot be null"):   # accepted as a shape: the call gets as far as the null check
        m.compact(None, None, shape, shape, 0)


def test_no_cpu_fallback_without_a_gpu():
    from DeepFlows.backend.backend_src.build.Release import CUDA_BACKEND as m
    if m.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="no usable CUDA device"):
        m.Array(16)
    lib = ctypes.CDLL(deepflows_b200.lib_path())
    p = ctypes.c_void_p()
    assert lib.dfb_malloc(ctypes.c_size_t(4), ctypes.byref(p)) != 0
    lib.dfb_last_error.restype = ctypes.c_char_p
    assert b"no CPU fallback" in lib.dfb_last_error()


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "deepflows_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), os.path.join(dirpath, f)
