"""Weight initialisers: host RNG (numpy.random) + one upload, same formulas and the same
fan computation as the reference (DeepFlows/nn/init.py:13-167), including its convention that
fan_in is read from shape[1] (so Linear weights stored (in,out) use `out` as fan_in)."""
import math
import warnings

import numpy as np

from ..autograd import no_grad
from ..tensor import Tensor


def _upload(tensor: Tensor, values):
    with no_grad():
        host = np.ascontiguousarray(values, dtype=np.float32)
        tensor.data = tensor.data.compact()
        tensor.data.device.from_numpy(host, tensor.data._handle)
    return tensor


def _no_grad_uniform_(tensor, a, b):
    return _upload(tensor, np.random.uniform(a, b, tensor.shape))


def _no_grad_normal_(tensor, mean, std):
    return _upload(tensor, np.random.normal(mean, std, size=tensor.shape))


def _no_grad_fill_(tensor, val):
    with no_grad():
        tensor.data.fill(val)
    return tensor


def calculate_gain(nonlinearity, param=None):
    linear_fns = ["linear", "conv1d", "conv2d", "conv3d", "conv_transpose1d", "conv_transpose2d", "conv_transpose3d"]
    if nonlinearity in linear_fns or nonlinearity == "sigmoid":
        return 1
    if nonlinearity == "tanh":
        return 5.0 / 3
    if nonlinearity == "relu":
        return math.sqrt(2.0)
    if nonlinearity == "leaky_relu":
        if param is None:
            slope = 0.01
        elif not isinstance(param, bool) and isinstance(param, (int, float)):
            slope = param
        else:
            raise ValueError("negative_slope {} not a valid number".format(param))
        return math.sqrt(2.0 / (1 + slope ** 2))
    if nonlinearity == "selu":
        return 3.0 / 4
    raise ValueError("Unsupported nonlinearity {}".format(nonlinearity))


def _calculate_fan_in_and_fan_out(tensor: Tensor):
    if tensor.ndim < 2:
        raise ValueError("Fan in and fan out can not be computed for tensor with fewer than 2 dimensions")
    receptive = 1
    for s in tensor.shape[2:]:
        receptive *= s
    return tensor.shape[1] * receptive, tensor.shape[0] * receptive


def normal_(tensor: Tensor, mean: float = 0., std: float = 1.) -> Tensor:
    return _no_grad_normal_(tensor, mean, std)


def uniform_(tensor: Tensor, low: float = 0., high: float = 1.0) -> Tensor:
    return _no_grad_uniform_(tensor, low, high)


def fill_(tensor: Tensor, val: float) -> Tensor:
    return _no_grad_fill_(tensor, val)


def zeros_(tensor: Tensor) -> Tensor:
    return _no_grad_fill_(tensor, 0.)


def ones_(tensor) -> Tensor:
    return _no_grad_fill_(tensor, 1.)


def xavier_uniform_(tensor: Tensor, gain: float = 1.0) -> Tensor:
    fan_in, fan_out = _calculate_fan_in_and_fan_out(tensor)
    bound = gain * math.sqrt(6. / (fan_in + fan_out))
    return _no_grad_uniform_(tensor, -bound, bound)


def xavier_normal_(tensor: Tensor, gain: float = 1.0) -> Tensor:
    fan_in, fan_out = _calculate_fan_in_and_fan_out(tensor)
    return _no_grad_normal_(tensor, 0., gain * math.sqrt(2.0 / (fan_in + fan_out)))


def _calculate_correct_fan(tensor, mode):
    mode = mode.lower()
    if mode not in ("fan_in", "fan_out"):
        raise ValueError("Mode {} not supported, please use one of fan_in, fan_out".format(mode))
    fan_in, fan_out = _calculate_fan_in_and_fan_out(tensor)
    return fan_in if mode == "fan_in" else fan_out


def _kaiming_std(tensor, a, mode, nonlinearity):
    return calculate_gain(nonlinearity, a) / math.sqrt(_calculate_correct_fan(tensor, mode))


def kaiming_uniform_(tensor: Tensor, a: float = 0, mode: str = "fan_in", nonlinearity: str = "relu"):
    # the reference's default here is "relu" (init.py:140-153), not torch's "leaky_relu": Conv2d / Linear call it with
    # a=sqrt(5) and get gain sqrt(2) (kaiming_normal_ below does default to "leaky_relu", init.py:156-168)
    if 0 in tensor.shape:
        warnings.warn("Initializing zero-element tensors is a no-op")
        return tensor
    bound = math.sqrt(3.0) * _kaiming_std(tensor, a, mode, nonlinearity)
    return _no_grad_uniform_(tensor, -bound, bound)


def kaiming_normal_(tensor: Tensor, a: float = 0, mode: str = "fan_in", nonlinearity: str = "leaky_relu"):
    if 0 in tensor.shape:
        warnings.warn("Initializing zero-element tensors is a no-op")
        return tensor
    return _no_grad_normal_(tensor, 0, _kaiming_std(tensor, a, mode, nonlinearity))
