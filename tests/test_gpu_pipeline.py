"""Per-batch preparation on the device (dfb_augment_batch, dfb_onehot_smooth): bit-exact against
the reference script's outputs (tests/golden/pipeline.npz) and, at the training batch size, against the oracle."""
import numpy as np
import pytest

from conftest import golden
from oracle import numpy_ops as ops

pytestmark = pytest.mark.gpu
F32 = np.float32


def _dev_tensor(a, dev):
    from DeepFlows import backend_api
    from DeepFlows.tensor import Tensor
    return Tensor(backend_api.Btensor(np.ascontiguousarray(a, dtype=F32), device=dev))


@pytest.mark.parametrize("case", ["late", "early_a", "early_b", "cifar"])
def test_device_augment_matches_reference(cuda_device, case):
    from DeepFlows.utils.data import BatchAugment
    g = golden("pipeline")
    x = g[case + "_x"]
    seed, epoch, num_epochs = (int(v) for v in g[case + "_meta"])
    aug = BatchAugment(pad=4)
    np.random.seed(seed)
    table = aug.draw(x.shape[0], x.shape[2], x.shape[3], epoch, num_epochs)
    launches = cuda_device.launch_count()
    y = aug(_dev_tensor(x, cuda_device), _dev_tensor(table, cuda_device))
    assert cuda_device.launch_count() == launches + 1          # one kernel, nothing on the host
    assert np.array_equal(y.numpy(), g[case + "_y"])
    y = aug(_dev_tensor(x, cuda_device), table)                # a host table is uploaded
    assert np.array_equal(y.numpy(), g[case + "_y"])


@pytest.mark.parametrize("shape,pad,clip", [((256, 3, 32, 32), 4, (-1.0, 1.0)), ((33, 1, 28, 28), 2, None),
                                            ((7, 5, 9, 13), 8, (-0.5, 0.25)), ((3, 2, 5, 5), 0, (-1.0, 1.0))])
def test_device_augment_matches_host_path(cuda_device, shape, pad, clip):
    """Full training batch and odd shapes; every batch erased (erase_p = 1), NaN / -0.0 pass through like numpy."""
    from DeepFlows.utils.data import BatchAugment
    rng = np.random.RandomState(shape[0])
    x = (rng.randn(*shape) * 1.5).astype(F32)
    x.reshape(-1)[::97] = -0.0
    x.reshape(-1)[5::211] = np.nan
    aug = BatchAugment(pad=pad, erase_p=1.0, erase_frac=(0.1, 0.6), clip=clip)
    np.random.seed(3)
    table = aug.draw(shape[0], shape[2], shape[3], epoch=0, num_epochs=20)
    assert table[:, 5].all()
    want = aug.apply_host(x, table)
    got = aug(_dev_tensor(x, cuda_device), _dev_tensor(table, cuda_device)).numpy()
    assert np.array_equal(got, want, equal_nan=True)
    ok = ~np.isnan(want)
    assert np.array_equal(np.signbit(got[ok]), np.signbit(want[ok]))   # -0.0 survives the clip, erased pixels are +0.0
    if pad == 4 and clip == (-1.0, 1.0):                        # the reference's configuration: oracle, same draws
        clean = np.nan_to_num(x, nan=0.25)
        np.random.seed(11)
        want = ops.augment_batch(clean, 2, 20)
        np.random.seed(11)
        table = BatchAugment(pad=4).draw(shape[0], shape[2], shape[3], 2, 20)
        got = BatchAugment(pad=4)(_dev_tensor(clean, cuda_device), table).numpy()
        assert np.array_equal(got, want)


def test_device_augment_rejects_bad_arguments(cuda_device):
    from DeepFlows.utils.data import BatchAugment
    x = _dev_tensor(np.zeros((2, 1, 4, 4), F32), cuda_device)
    with pytest.raises(ValueError):
        BatchAugment(pad=4)(x, np.zeros((2, 8), F32))           # numpy's reflect needs pad < size
    with pytest.raises(ValueError):
        BatchAugment(pad=1)(x, np.zeros((3, 8), F32))           # one table row per sample
    xb = x.data
    with pytest.raises(ValueError):
        cuda_device.augment_batch(xb._handle, xb._handle, xb._handle, 2, 1, 4, 4, 1, True, -1.0, 1.0)  # in place


def test_device_one_hot_label_smoothing(cuda_device):
    from DeepFlows.utils.data import smooth_one_hot
    g = golden("pipeline")
    classes, eps = int(g["smooth_meta"][0]), float(g["smooth_meta"][1])
    labels = _dev_tensor(g["labels"].astype(F32), cuda_device)
    got = smooth_one_hot(labels, classes, eps)
    assert got.shape == (len(g["labels"]), classes) and np.array_equal(got.numpy(), g["smoothed"])
    assert np.array_equal(smooth_one_hot(labels, classes).numpy(), g["onehot"])
    assert np.array_equal(smooth_one_hot(g["labels"], classes, eps, device=cuda_device).numpy(), g["smoothed"])
    big = np.random.RandomState(0).randint(0, 1000, size=4096)
    assert np.array_equal(smooth_one_hot(big, 1000, 0.1, device=cuda_device).numpy(), ops.smooth_one_hot(big, 1000, 0.1))


def test_prefetched_batches_are_augmented_on_the_device(cuda_device):
    """The whole input side of a step: loader -> (x, labels, draws) -> DevicePrefetcher -> augment + targets on the
    device == the reference's host code on the same generator stream."""
    from DeepFlows.utils.data import BatchAugment, DevicePrefetcher, data_loader, smooth_one_hot
    rng = np.random.RandomState(4)
    X = rng.randn(40, 3, 16, 16).astype(F32)
    L = rng.randint(0, 10, 40).astype(F32)
    aug = BatchAugment(pad=4)
    np.random.seed(21)
    want = [(ops.augment_batch(X[i:i + 16], 0, 20), ops.smooth_one_hot(L[i:i + 16], 10, 0.05)) for i in range(0, 40, 16)]
    np.random.seed(21)
    loader = ((x, l, aug.draw(len(x), 16, 16, 0, 20)) for x, l in data_loader(X, L, batch_size=16))
    seen = 0
    for (x, l, table), (wx, wt) in zip(DevicePrefetcher(loader, device=cuda_device), want):
        assert np.array_equal(aug(x, table).numpy(), wx)
        assert np.array_equal(smooth_one_hot(l, 10, 0.05).numpy(), wt)
        seen += 1
    assert seen == 3
