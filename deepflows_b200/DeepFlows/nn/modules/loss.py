"""Loss modules: thin Module wrappers over the functional losses, same names and constructor as the reference's
(DeepFlows/nn/modules/loss.py:22-60: `Loss(reduction='mean' | 'sum')(input, target)`). The classes are generated from
one table, since each differs only in the functional it forwards to."""
from .module import Module
from .. import functional as F

_FUNCTIONALS = {
    "L1Loss": "l1_loss",
    "NLLLoss": "nll_loss",
    "MSELoss": "mse_loss",
    "BCELoss": "binary_cross_entropy",
    "CrossEntropyLoss": "cross_entropy",   # dense (one-hot / smoothed) target rows; fused softmax-CE kernel on cuda
}
__all__ = list(_FUNCTIONALS)


class _Loss(Module):
    functional = None  # name of the function in nn.functional

    def __init__(self, reduction: str = "mean") -> None:
        super().__init__()
        if reduction not in ("mean", "sum"):
            raise AssertionError("reduction must be 'mean' or 'sum'")
        self.reduction = reduction

    def forward(self, input, target):
        if self.functional is None:
            raise NotImplementedError
        return getattr(F, self.functional)(input, target, reduction=self.reduction)


for _name, _fn in _FUNCTIONALS.items():
    globals()[_name] = type(_name, (_Loss,), {"functional": _fn, "__module__": __name__,
                                              "__doc__": "reduction(F.%s(input, target))" % _fn})
del _name, _fn
