import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import deepflows_b200  # noqa: E402,F401  (puts the DeepFlows host package on sys.path)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200; run with `-m gpu` on the GPU box")


def golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def rel_err(got, want):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, (got.shape, want.shape)
    return float(np.abs(got - want).max() / max(np.abs(want).max(), 1e-30))


@pytest.fixture(scope="session")
def cpu_device():
    """The oracle's numpy device registered as DeepFlows' `cpu` (test infrastructure, not product)."""
    from oracle import numpy_device
    from DeepFlows import backend_api
    return backend_api.register_numpy_device(numpy_device)


@pytest.fixture(scope="session")
def cuda_device():
    from DeepFlows import backend_api
    dev = backend_api.cuda()
    assert dev.enabled(), "CUDA_BACKEND extension is not built"
    assert dev.device_count() > 0, "no CUDA device"
    return dev


@pytest.fixture(autouse=True)
def _fresh_graph():
    from DeepFlows import tensor, autograd, backend_api
    tensor.Graph.free_graph_all()
    autograd.set_grad_enabled(True)
    backend_api.set_dgrad_mode("exact")  # the package default; fixture-parity tests select "reference" themselves
    yield
    tensor.Graph.free_graph_all()
    autograd.set_grad_enabled(True)
    backend_api.set_dgrad_mode("exact")
