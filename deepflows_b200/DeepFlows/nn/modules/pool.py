"""Pooling layers (reference: DeepFlows/nn/modules/pool.py)."""
from .module import Module
from .. import functional as F


class _Pool(Module):
    def __init__(self, kernel_size: int, stride: int = 0, padding: int = 0) -> None:
        super().__init__()
        self.kernel_size, self.stride, self.padding = kernel_size, stride, padding

    def __repr__(self) -> str:
        return "{}(kernel_size={}, stride={}, padding={})".format(type(self).__name__, self.kernel_size, self.stride,
                                                                 self.padding)


class MaxPool1d(_Pool):
    def forward(self, x):
        return F.max_pool1d(x, self.kernel_size, self.stride, self.padding)


class AvgPool1d(_Pool):
    def forward(self, x):
        return F.avg_pool1d(x, self.kernel_size, self.stride, self.padding)


class MaxPool2d(_Pool):
    def forward(self, x):
        return F.max_pool2d(x, self.kernel_size, self.stride, self.padding)


class AvgPool2d(_Pool):
    def forward(self, x):
        return F.avg_pool2d(x, self.kernel_size, self.stride, self.padding)
