"""Adagrad (reference: DeepFlows/optim/adagrad.py). The reference keeps its state in host numpy arrays
(`np.zeros`, line 16) and therefore cannot run on a cuda device; here the state lives on the device
and the update is composed from BackendTensor ops (not a fused kernel: outside the benchmarked path)."""
from .optimier import Optimizer
from .. import backend_api


class Adagrad(Optimizer):
    def __init__(self, params, lr: float = 1e-2, weight_decay: float = 0.0, eps: float = 1e-10) -> None:
        super().__init__(params)
        self.lr, self.weight_decay, self.eps = lr, weight_decay, eps
        self.s = [backend_api.zeros_like(p.data) for p in self.params]

    def step(self):
        # data parallel: wait for the bucketed all-reduces on the compute stream and average (1/world_size)
        grad_scale = self._grad_scale()
        for i, p, g in self._active():
            if grad_scale != 1.0:
                g = g * grad_scale
            if self.weight_decay:
                g = g + p.data * self.weight_decay
            self.s[i] = self.s[i] + g * g
            p.data = p.data - g * self.lr / (self.s[i] + self.eps) ** 0.5
