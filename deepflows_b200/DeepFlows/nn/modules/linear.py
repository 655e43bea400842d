"""Linear layer: y = x @ W + b with W stored (in_features, out_features) and b (1, out_features)
(reference: DeepFlows/nn/modules/linear.py:10-67)."""
import math

from .module import Module
from ..parameter import Parameter
from .. import functional as F
from .. import init
from ...tensor import empty, Tensor
from ... import backend_api


class Linear(Module):
    __constants__ = ["in_features", "out_features"]

    def __init__(self, in_features: int, out_features: int, bias: bool = True, device="cuda", dtype="float32") -> None:
        super().__init__()
        kwargs = {"device": backend_api.Device(device), "dtype": dtype}
        self.in_features, self.out_features = in_features, out_features
        self.weight = Parameter(empty((in_features, out_features), **kwargs))
        if bias:
            self.bias = Parameter(empty((1, out_features), **kwargs))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self) -> None:
        init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            fan_in, _ = init._calculate_fan_in_and_fan_out(self.weight)
            bound = 1 / math.sqrt(fan_in) if fan_in > 0 else 0
            init.uniform_(self.bias, -bound, bound)

    def forward(self, input: Tensor) -> Tensor:
        return F.linear(input, self.weight, self.bias)

    def extra_repr(self) -> str:
        return "in_features={}, out_features={}, bias={}".format(self.in_features, self.out_features,
                                                                 self.bias is not None)

    def move(self, device):
        self.device = device
        self.weight = self.weight.to(device)
        if self.bias is not None:
            self.bias = self.bias.to(device)
