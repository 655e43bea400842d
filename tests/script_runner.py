"""Runs one of the reference's own training scripts (test/*_cuda.py) UNCHANGED for a few steps and records the loss of
every training step - the "scripts run unchanged" acceptance of BASELINE.json's north_star.

The script source is read from $DEEPFLOWS_REFERENCE/test (default /root/reference/test; never copied into this repo) and
exec'd as `__main__`. What the harness supplies around it ("portability hazards"):
  * plotting / dataset packages that are absent or need the network: matplotlib, seaborn, PIL, sklearn.datasets.fetch_openml
    (stub modules in sys.modules);
  * the datasets (blobs stripped from the reference): `open()` of an MNIST idx file or a CIFAR-10 pickle - whatever the
    directory, including the scripts' hard-coded Windows paths - returns a small seeded synthetic file;
  * a stop after `steps` optimizer steps (the scripts train for 10-50 epochs): `Optimizer.step` of the package under test is
    wrapped to count, and the loss the script computed for that step (the last `CrossEntropyLoss` result) is recorded.
Nothing of the model / loss / optimizer / loop code is touched.

Which package the script's `from DeepFlows import ...` resolves to, and what `device='cuda'` means:
  --package ours      --device cuda     deepflows_b200's host package on libdfb200 (needs a GPU)
  --package ours      --device oracle   the same host package with the oracle's numpy device standing in for 'cuda'
  --package reference --device oracle   the reference's own host package on its numpy device with the two setitem functions
                                        restated from its CUDA kernels - the definition of the oracle

    python tests/script_runner.py --script CNN_MNIST_cuda.py --package ours --device oracle --steps 2 --out /tmp/a.npz
"""
import argparse
import builtins
import io
import os
import pickle
import struct
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("DEEPFLOWS_REFERENCE", "/root/reference")


class _Stop(Exception):
    pass


class _Anything(types.ModuleType):
    """A module whose every attribute is a callable returning another _Anything (plt.figure(...).add_subplot(...) ...)."""

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Anything(name)

    def __call__(self, *a, **k):
        return _Anything("call")

    def __iter__(self):
        return iter(())


def _stub_modules(samples):
    for name in ("matplotlib", "matplotlib.pyplot", "seaborn", "PIL", "PIL.Image", "nvtx"):
        sys.modules[name] = _Anything(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["PIL"].Image = sys.modules["PIL.Image"]
    try:  # fetch_openml needs the network: a seeded synthetic MNIST-shaped frame instead
        import sklearn.datasets as skd
        import pandas as pd

        def fetch_openml(name, version=1, return_X_y=True, **kw):
            rng = np.random.RandomState(7)
            x = pd.DataFrame(rng.randint(0, 256, (max(5000, samples), 784)).astype(np.float64))
            y = pd.Series(rng.randint(0, 10, max(5000, samples)).astype(np.int64))   # (the real set has string categories)
            return x, y
        skd.fetch_openml = fetch_openml
    except ImportError:
        pass


def _idx_bytes(arr):
    arr = np.ascontiguousarray(arr, dtype=np.uint8)
    head = bytes([0, 0, 8, arr.ndim]) + b"".join(struct.pack(">I", s) for s in arr.shape)
    return head + arr.tobytes()


def _synthetic_file(path, samples):
    """Bytes of a small seeded dataset file for the basenames the scripts open, else None."""
    base = os.path.basename(str(path).replace("\\", "/"))
    rng = np.random.RandomState(sum(base.encode()))
    if base in ("train-images-idx3-ubyte", "t10k-images-idx3-ubyte"):
        return _idx_bytes(rng.randint(0, 256, (samples, 28, 28)))
    if base in ("train-labels-idx1-ubyte", "t10k-labels-idx1-ubyte"):
        return _idx_bytes(rng.randint(0, 10, (samples,)))
    if base.startswith("data_batch_") or base == "test_batch":
        n = max(1, samples // 5) if base.startswith("data_batch_") else samples
        return pickle.dumps({"data": rng.randint(0, 256, (n, 3072)).astype(np.uint8), "labels": rng.randint(0, 10, n).tolist()})
    return None


def _import_package(package, device):
    """Puts the requested DeepFlows package on sys.path / in sys.modules and makes Device('cuda') the requested device."""
    sys.path.insert(0, ROOT)
    if package == "reference":
        os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
        sys.dont_write_bytecode = True
        sys.path.insert(0, REF)
        import DeepFlows
        assert os.path.realpath(DeepFlows.__file__).startswith(os.path.realpath(REF)), DeepFlows.__file__
        from DeepFlows.backend import backend_tensor as rbt
        from oracle import numpy_ops as ops
        assert device == "oracle", "the reference ships no Linux CUDA extension"
        dev = rbt.cpu_numpy()
        dev.mod.ewise_setitem = lambda a, out, shape, strides, offset: ops.ewise_setitem(a, out, shape, strides, offset)
        dev.mod.scalar_setitem = lambda size, value, out, shape, strides, offset: ops.scalar_setitem(size, value, out, shape, strides, offset)
        as_cuda = rbt.BackendDevice("cuda", dev.mod)
        rbt.cuda = lambda: as_cuda       # `Device('cuda')` -> all_devices() -> cuda()
        rbt.cpu_numpy = lambda: dev
        return DeepFlows
    import deepflows_b200  # noqa: F401
    import DeepFlows
    from DeepFlows.backend import backend_tensor as bt
    if device == "oracle":
        from oracle import numpy_device
        bt._cuda_device = bt.BackendDevice("cuda", numpy_device)   # the oracle's numpy device stands in for 'cuda'
        bt.set_dgrad_mode("reference")   # the reference's own conv backward, like the package it is compared with
    else:
        assert bt.cuda().enabled(), "CUDA_BACKEND is not built"
        bt.set_dgrad_mode(os.environ.get("DEEPFLOWS_DGRAD", "reference"))
    return DeepFlows


def run(script, package, device, steps, samples=96, seed=0):
    DeepFlows = _import_package(package, device)
    _stub_modules(samples)
    from DeepFlows import nn, optim
    losses, last = [], {}

    # the loss of each step: remember what CrossEntropyLoss returned, read it when the optimizer steps
    ce = nn.CrossEntropyLoss
    ce_forward = ce.forward

    def forward(self, *a, **k):
        out = ce_forward(self, *a, **k)
        last["loss"] = out
        return out
    ce.forward = forward
    seen = set()
    for name in dir(optim):
        cls = getattr(optim, name)
        if isinstance(cls, type) and hasattr(cls, "step") and cls.__dict__.get("step") is not None and cls not in seen:
            seen.add(cls)

            def make(orig):
                def step(self, *a, **k):
                    r = orig(self, *a, **k)
                    if "loss" in last:
                        losses.append(float(last["loss"].data.numpy().reshape(-1)[0]))
                    if len(losses) >= steps:
                        raise _Stop()
                    return r
                return step
            cls.step = make(cls.__dict__["step"])

    path = os.path.join(REF, "test", script)
    src = open(path, encoding="utf-8").read()
    real_open = builtins.open

    def fake_open(file, mode="r", *a, **k):
        data = _synthetic_file(file, samples) if "b" in mode else None
        return io.BytesIO(data) if data is not None else real_open(file, mode, *a, **k)

    # `__file__` stays the script's real location (its sys.path.insert of the parent directory is harmless: DeepFlows is
    # already imported); `open` is resolved in the script's globals first
    ns = {"__name__": "__main__", "__file__": path, "open": fake_open}
    np.random.seed(seed)
    import random
    random.seed(seed)
    builtins.open, saved_stdout = fake_open, sys.stdout
    sys.stdout = io.StringIO()  # the scripts print per batch
    if not hasattr(sys.stdout, "reconfigure"):
        sys.stdout.reconfigure = lambda **k: None
    try:
        exec(compile(src, path, "exec"), ns)
    except _Stop:
        pass
    finally:
        builtins.open, sys.stdout = real_open, saved_stdout
    return np.array(losses, dtype=np.float64)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--script", required=True)
    ap.add_argument("--package", default="ours", choices=["ours", "reference"])
    ap.add_argument("--device", default="oracle", choices=["oracle", "cuda"])
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--samples", type=int, default=96)
    ap.add_argument("--out", required=True)
    a = ap.parse_args()
    res = run(a.script, a.package, a.device, a.steps, a.samples)
    np.savez(a.out, losses=res)
    print("%s %s/%s losses %s" % (a.script, a.package, a.device, res))
