"""SURVEY 8f rank 4 ops on a B200: general pooling, 1-d ops, activation modules, device-state Adagrad / Adadelta."""
import pytest

import ops_f4

pytestmark = pytest.mark.gpu


def test_general_pooling(cuda_device):
    ops_f4.check_general_pooling(cuda_device)


def test_1d_ops(cuda_device):
    from DeepFlows import backend_api
    backend_api.set_precision("fp32")
    ops_f4.check_1d_ops(cuda_device)


def test_activation_modules(cuda_device):
    ops_f4.check_activations(cuda_device)


def test_adagrad_adadelta(cuda_device):
    ops_f4.check_adagrad_adadelta(cuda_device)


def test_device_dropout(cuda_device):
    import numpy as np
    m = cuda_device.mod

    def mask_of(n, keep, seed, step):
        hm, hs = m.Array(n), m.Array(2)
        m.from_numpy(np.array([seed, step], np.float32), hs)
        m.dropout_mask(hm, n, keep, hs)
        return m.to_numpy(hm, [n], [1], 0)
    ops_f4.check_device_dropout(cuda_device, mask_of)
