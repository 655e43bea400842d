"""Builds the native pieces of deepflows_b200 in-tree, for sm_100a only.

  deepflows_b200/lib/libdfb200.so
      the C-ABI backend (include/dfb200.h) from csrc/*.cu
  deepflows_b200/DeepFlows/backend/backend_src/build/Release/CUDA_BACKEND.<ext>.so
      the pybind shim, at the dotted path the reference imports
      (reference: DeepFlows/backend/backend_tensor.py:57)

The reference builds its one .cu with CMake for sm_52 into a Windows .pyd
(DeepFlows/backend/backend_src/CMakeLists.txt:20-46); here it is plain nvcc/g++ invocations so the
artefacts land next to the sources and travel with the tree.

Usage: python -m deepflows_b200.build [--force] [--verbose]
"""
import os
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "lib", "obj")
LIB = os.path.join(HERE, "lib", "libdfb200.so")
SHIM_DIR = os.path.join(HERE, "DeepFlows", "backend", "backend_src", "build", "Release")
SHIM = os.path.join(SHIM_DIR, "CUDA_BACKEND" + sysconfig.get_config_var("EXT_SUFFIX"))

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC,-fvisibility=hidden", "-Xptxas", "-v",
]
CU_SOURCES = ["runtime.cu", "ewise.cu", "gemm_simt.cu", "gemm_tc.cu", "conv_direct.cu", "gemm.cu", "nn_ops.cu", "data_ops.cu", "optim.cu", "comm.cu", "peer.cu"]
HEADERS = [os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "kernels.cuh"), os.path.join(ROOT, "include", "dfb200.h")]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _run(cmd, verbose, log=None):
    if verbose:
        print("+", " ".join(cmd), flush=True)
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if log is not None:
        with open(log, "w") as f:
            f.write(" ".join(cmd) + "\n" + p.stdout)
    if p.returncode != 0:
        raise RuntimeError("build step failed:\n  %s\n%s" % (" ".join(cmd), p.stdout))
    return p.stdout


def build_all(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(SHIM_DIR, exist_ok=True)
    extra_headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    deps_common = sorted(set(HEADERS + extra_headers))

    def compile_one(src):
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        if force or _newer(o, [s] + deps_common):
            _run([NVCC] + NVCC_FLAGS + ["-c", s, "-o", o], verbose, log=o + ".log")
            return o, True
        return o, False

    with ThreadPoolExecutor(max_workers=min(8, len(CU_SOURCES))) as ex:
        results = list(ex.map(compile_one, CU_SOURCES))
    objs = [o for o, _ in results]
    if force or any(ch for _, ch in results) or not os.path.exists(LIB):
        _run([NVCC, "-shared", "-o", LIB] + objs + ["-cudart", "static", "-ldl"], verbose)

    shim_src = os.path.join(CSRC, "pybind_shim.cpp")
    if force or _newer(SHIM, [shim_src, LIB] + HEADERS):
        import pybind11

        inc = ["-I" + pybind11.get_include(), "-I" + sysconfig.get_paths()["include"]]
        rel = os.path.relpath(os.path.dirname(LIB), SHIM_DIR)
        _run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-fvisibility=hidden"] + inc +
             [shim_src, "-o", SHIM, "-L" + os.path.dirname(LIB), "-ldfb200",
              "-Wl,-rpath,$ORIGIN/" + rel], verbose)
    return LIB, SHIM


if __name__ == "__main__":
    lib, shim = build_all(force="--force" in sys.argv, verbose="--verbose" in sys.argv or "-v" in sys.argv)
    print(lib)
    print(shim)
