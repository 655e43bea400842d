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
