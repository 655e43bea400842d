"""Per-batch preparation without a GPU: the oracle's restatement and the host package's
`BatchAugment` / `smooth_one_hot` against tests/golden/pipeline.npz, which oracle/make_golden_pipeline.py took from
the reference's own training script (test/ResNet_CIFAR10_cuda.py:129-148, 181-183)."""
import numpy as np
import pytest

from conftest import golden
from oracle import numpy_ops as ops

CASES = ["late", "early_a", "early_b", "cifar"]


@pytest.mark.parametrize("case", CASES)
def test_oracle_augment_matches_reference(case):
    g = golden("pipeline")
    seed, epoch, num_epochs = (int(v) for v in g[case + "_meta"])
    np.random.seed(seed)
    got = ops.augment_batch(g[case + "_x"], epoch, num_epochs)
    assert got.dtype == np.float32 and np.array_equal(got, g[case + "_y"])


@pytest.mark.parametrize("case", CASES)
def test_host_draws_follow_the_reference_generator_order(case):
    """BatchAugment.draw consumes numpy's generator like the reference function, so a seeded run augments the same
    way; apply_host on that table reproduces the reference output bit for bit, and the generator is left in the same
    state (the NEXT batch would match as well)."""
    from DeepFlows.utils.data import BatchAugment
    g = golden("pipeline")
    x = g[case + "_x"]
    seed, epoch, num_epochs = (int(v) for v in g[case + "_meta"])
    aug = BatchAugment(pad=4)
    np.random.seed(seed)
    table = aug.draw(x.shape[0], x.shape[2], x.shape[3], epoch, num_epochs)
    after_mine = np.random.rand()
    assert table.shape == (x.shape[0], 8) and table.dtype == np.float32
    assert np.array_equal(aug.apply_host(x, table), g[case + "_y"])
    assert np.array_equal(aug(x, table), g[case + "_y"])          # host arrays take the numpy path
    np.random.seed(seed)
    ops.augment_batch(x, epoch, num_epochs)
    assert after_mine == np.random.rand()
    erased = bool(table[:, 5].any())
    assert erased == (case != "late")


def test_table_semantics_by_hand():
    from DeepFlows.utils.data import BatchAugment
    x = np.arange(2 * 1 * 3 * 4, dtype=np.float32).reshape(2, 1, 3, 4) / 100
    aug = BatchAugment(pad=1, clip=None)
    table = np.zeros((2, 8), np.float32)
    table[0, :3] = (1, 1, 0)            # centre crop: identity
    table[1, :3] = (0, 2, 1)            # one row up, one column right, mirrored
    table[1, 3:7] = (2, 0, 1, 2)        # zero the last row's first two pixels
    y = aug.apply_host(x, table)
    assert np.array_equal(y[0], x[0])
    padded = np.pad(x[1, 0], 1, mode="reflect")
    want = padded[0:3, 2:6][:, ::-1].copy()
    want[2, 0:2] = 0
    assert np.array_equal(y[1, 0], want)


def test_oracle_and_host_smooth_one_hot():
    from DeepFlows.utils.data import smooth_one_hot
    g = golden("pipeline")
    classes, eps = int(g["smooth_meta"][0]), float(g["smooth_meta"][1])
    assert np.array_equal(ops.smooth_one_hot(g["labels"], classes, eps), g["smoothed"])
    assert np.array_equal(ops.smooth_one_hot(g["labels"], classes, 0.0), g["onehot"])
    got = smooth_one_hot(g["labels"], classes, eps)            # no device: numpy in, numpy out
    assert isinstance(got, np.ndarray) and got.dtype == np.float32 and np.array_equal(got, g["smoothed"])
    assert np.array_equal(smooth_one_hot(g["labels"], classes), g["onehot"])
