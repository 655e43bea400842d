"""deepflows_b200: a B200 (sm_100a) compute backend for the DeepFlows framework.

  csrc/ + lib/libdfb200.so   hand-written CUDA kernels behind the C ABI of include/dfb200.h
  DeepFlows/                 host-side mirror of the reference's Python package (same import names);
                             importing `deepflows_b200` makes it importable as top-level `DeepFlows`
  build.py                   in-tree build of the library and the `CUDA_BACKEND` pybind shim
"""
import os as _os
import sys as _sys

_here = _os.path.dirname(_os.path.abspath(__file__))
if _here not in _sys.path:
    _sys.path.insert(0, _here)

__version__ = "0.1.0"


def lib_path():
    return _os.path.join(_here, "lib", "libdfb200.so")
