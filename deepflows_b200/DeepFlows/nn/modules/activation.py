"""Activation modules (reference: DeepFlows/nn/modules/activation.py)."""
from typing import Optional

from .module import Module
from .. import functional as F
from ...tensor import Tensor

__all__ = ["ReLU", "Sigmoid", "Tanh", "LeakyReLU", "Softmax", "LogSoftmax"]


class ReLU(Module):
    def forward(self, input: Tensor) -> Tensor:
        return F.relu(input)


class Sigmoid(Module):
    def forward(self, input: Tensor) -> Tensor:
        return F.sigmoid(input)


class Tanh(Module):
    def forward(self, input: Tensor) -> Tensor:
        return F.tanh(input)


class LeakyReLU(Module):
    def __init__(self, negative_slope: float = 1e-2) -> None:
        super().__init__()
        self.negative_slope = negative_slope

    def forward(self, input: Tensor) -> Tensor:
        return F.leaky_relu(input, self.negative_slope)

    def extra_repr(self) -> str:
        return "negative_slope={}".format(self.negative_slope)


class Softmax(Module):
    def __init__(self, dim: Optional[int] = None) -> None:
        super().__init__()
        self.dim = dim

    def forward(self, input: Tensor) -> Tensor:
        return F.softmax(input, self.dim)

    def extra_repr(self) -> str:
        return "dim={}".format(self.dim)


class LogSoftmax(Module):
    def __init__(self, dim: Optional[int] = None) -> None:
        super().__init__()
        self.dim = dim

    def forward(self, input: Tensor) -> Tensor:
        return F.log_softmax(input, self.dim)

    def extra_repr(self) -> str:
        return "dim={}".format(self.dim)
