"""Loss modules (reference: DeepFlows/nn/modules/loss.py)."""
from .module import Module
from .. import functional as F
from ...tensor import Tensor

__all__ = ["L1Loss", "NLLLoss", "MSELoss", "BCELoss", "CrossEntropyLoss"]


class _Loss(Module):
    def __init__(self, reduction: str = "mean") -> None:
        super().__init__()
        assert reduction in {"mean", "sum"}
        self.reduction = reduction


class L1Loss(_Loss):
    def forward(self, input: Tensor, target: Tensor) -> Tensor:
        return F.l1_loss(input, target, reduction=self.reduction)


class NLLLoss(_Loss):
    def forward(self, input: Tensor, target: Tensor) -> Tensor:
        return F.nll_loss(input, target, reduction=self.reduction)


class MSELoss(_Loss):
    def forward(self, input: Tensor, target: Tensor):
        return F.mse_loss(input, target, reduction=self.reduction)


class BCELoss(_Loss):
    def forward(self, input: Tensor, target: Tensor):
        return F.binary_cross_entropy(input, target, reduction=self.reduction)


class CrossEntropyLoss(_Loss):
    def forward(self, input: Tensor, target: Tensor) -> Tensor:
        return F.cross_entropy(input, target, reduction=self.reduction)
