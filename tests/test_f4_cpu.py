"""SURVEY 8f rank 4 ops on the oracle numpy device (host logic); the GPU tier runs the same checks on libdfb200."""
import ops_f4


def test_general_pooling(cpu_device):
    ops_f4.check_general_pooling(cpu_device)


def test_1d_ops(cpu_device):
    ops_f4.check_1d_ops(cpu_device)


def test_activation_modules(cpu_device):
    ops_f4.check_activations(cpu_device)


def test_adagrad_adadelta(cpu_device):
    ops_f4.check_adagrad_adadelta(cpu_device)


def test_device_dropout(cpu_device):
    import numpy as np
    from oracle import numpy_device as nd

    def mask_of(n, keep, seed, step):
        m, st = nd.Array(n), nd.Array(2)
        st.buf[:] = [seed, step]
        nd.dropout_mask(m, n, keep, st)
        return m.buf.copy()
    ops_f4.check_device_dropout(cpu_device, mask_of)
