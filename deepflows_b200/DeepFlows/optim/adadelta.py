"""Adadelta (reference: DeepFlows/optim/adadelta.py; host-numpy state there, device state here)."""
from .optimier import Optimizer
from .. import backend_api


class Adadelta(Optimizer):
    def __init__(self, params, lr: float = 1.0, rho: float = 0.9, weight_decay: float = 0.0, eps: float = 1e-6) -> None:
        super().__init__(params)
        self.lr, self.rho, self.weight_decay, self.eps = lr, rho, weight_decay, eps
        self.s = [backend_api.zeros_like(p.data) for p in self.params]
        self.delta = [backend_api.zeros_like(p.data) for p in self.params]

    def step(self):
        # data parallel: wait for the bucketed all-reduces on the compute stream and average (1/world_size)
        grad_scale = self._grad_scale()
        for i, p, g in self._active():
            if grad_scale != 1.0:
                g = g * grad_scale
            if self.weight_decay:
                g = g + p.data * self.weight_decay
            self.s[i] = self.s[i] * self.rho + g * g * (1 - self.rho)
            upd = g * ((self.delta[i] + self.eps) ** 0.5) / ((self.s[i] + self.eps) ** 0.5)
            self.delta[i] = self.delta[i] * self.rho + upd * upd * (1 - self.rho)
            p.data = p.data - upd  # the reference applies the adjusted gradient without lr
