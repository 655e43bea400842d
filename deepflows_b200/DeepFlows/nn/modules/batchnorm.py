"""BatchNorm2d (reference: DeepFlows/nn/modules/batchnorm.py:8-64).

Parameters and running statistics have shape (1, C, 1, 1). Training uses the batch mean and the
biased variance and folds the *biased* variance into running_var (reference lines 44-46, SURVEY
Q7). The running statistics are additionally registered as buffers - a superset of the reference,
where they are plain attributes and therefore invisible to checkpoints.
"""
from .module import Module
from ..parameter import Parameter
from .. import functional as F
from ...tensor import Tensor
from ... import tensor
from ... import backend_api


class BatchNorm2d(Module):
    def __init__(self, num_features: int, eps: float = 1e-5, momentum: float = 0.1, affine: bool = True,
                 track_running_stats: bool = True, device: str = "cuda", dtype=None) -> None:
        super().__init__()
        kwargs = {"device": backend_api.Device(device), "dtype": dtype}
        self.num_features, self.eps, self.momentum = num_features, eps, momentum
        self.affine, self.track_running_stats = affine, track_running_stats
        shape = (1, num_features, 1, 1)
        if affine:
            self.weight = Parameter(tensor.ones(shape, **kwargs))
            self.bias = Parameter(tensor.zeros(shape, **kwargs))
        else:
            self.weight = None
            self.bias = None
        if track_running_stats:
            self.register_buffer("running_mean", tensor.zeros(shape, **kwargs))
            self.register_buffer("running_var", tensor.ones(shape, **kwargs))
        else:
            self.running_mean = None
            self.running_var = None

    def forward(self, x: Tensor) -> Tensor:
        return F.batch_norm(x, self.weight, self.bias, self.running_mean, self.running_var, self.training,
                            self.momentum, self.eps)

    def __repr__(self) -> str:
        return "{}(num_features={}, eps={}, momentum={}, affine={}, track_running_stats={})".format(
            type(self).__name__, self.num_features, self.eps, self.momentum, self.affine, self.track_running_stats)
