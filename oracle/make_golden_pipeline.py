"""Generates tests/golden/pipeline.npz from the REFERENCE's own training script and pins the oracle's restatement
(oracle/numpy_ops.py: augment_batch, smooth_one_hot) against it.

Runs only in the build container (it imports /root/reference/test/ResNet_CIFAR10_cuda.py, which does not exist on
the GPU box); the fixture is committed and travels. Usage:

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_pipeline.py

The script's `augment_batch` (lines 129-148) is called as is. Its module imports matplotlib, which this image does
not have and the function does not use: an empty stand-in module is registered before the import. The smoothed
one-hot targets are three inline statements of `train_resnet` (lines 159-161, 181-183); they are executed here with
the same sklearn encoder the script builds.
"""
import importlib.util
import os
import sys
import types

os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("DEEPFLOWS_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from oracle import numpy_ops as ops  # noqa: E402


def load_reference_script():
    for name in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    path = os.path.join(REF, "test", "ResNet_CIFAR10_cuda.py")
    spec = importlib.util.spec_from_file_location("ref_resnet_cifar10_cuda", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)  # definitions only: the training run sits behind __main__
    return mod


def main():
    ref = load_reference_script()
    from sklearn.preprocessing import OneHotEncoder
    out = {}
    # (name, shape, epoch, num_epochs): the erase branch is only reachable while epoch < num_epochs - 5
    cases = [("late", (6, 3, 12, 10), 18, 20), ("early_a", (6, 3, 12, 10), 0, 20), ("early_b", (5, 2, 9, 16), 3, 20),
             ("cifar", (4, 3, 32, 32), 1, 20)]
    erased = 0
    for name, shape, epoch, num_epochs in cases:
        data_rng = np.random.RandomState(len(name) * 7 + shape[0])
        x = (data_rng.randn(*shape) * 1.5).astype(np.float32)  # values beyond [-1, 1] so the clip matters
        want_erase = name.startswith("early") or name == "cifar"
        seed = 0
        while True:  # find a seed whose batch takes (early_*) / skips (late) the erase branch
            np.random.seed(seed)
            got = ref.augment_batch(x, epoch, num_epochs)
            has_zero_block = bool((got == 0.0).any())
            if has_zero_block == want_erase:
                break
            seed += 1
        np.random.seed(seed)
        mine = ops.augment_batch(x, epoch, num_epochs)
        assert mine.dtype == got.dtype and np.array_equal(mine, got), name
        erased += has_zero_block
        out[name + "_x"], out[name + "_y"] = x, got
        out[name + "_meta"] = np.array([seed, epoch, num_epochs], dtype=np.int64)
        print("augment_batch %-8s shape %-16s seed %3d erase %-5s oracle == reference" % (name, shape, seed, has_zero_block))
    assert erased >= 2

    num_classes, eps = 10, 0.05
    encoder = OneHotEncoder(sparse_output=False)
    encoder.fit(np.arange(num_classes).reshape(-1, 1))
    labels = np.random.RandomState(5).randint(0, num_classes, size=37)
    onehot = encoder.transform(labels.reshape(-1, 1)).astype(np.float32)
    smoothed = onehot * (1 - eps) + eps / num_classes
    assert smoothed.dtype == np.float32
    mine = ops.smooth_one_hot(labels, num_classes, eps)
    assert mine.dtype == np.float32 and np.array_equal(mine, smoothed)
    assert np.array_equal(ops.smooth_one_hot(labels, num_classes, 0.0), onehot)
    out["labels"], out["smoothed"], out["onehot"] = labels.astype(np.int64), smoothed, onehot
    out["smooth_meta"] = np.array([num_classes, eps], dtype=np.float64)
    print("smooth_one_hot oracle == reference statements")

    path = os.path.join(ROOT, "tests", "golden", "pipeline.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
