"""SGD with momentum / Nesterov / L2 decay (reference: DeepFlows/optim/sgd.py:8-24):
g = grad + p*wd (always); v = v*momentum + g; update = g + momentum*v (nesterov) or v; p -= lr*update.
One fused launch for all parameters."""
from typing import List

from .optimier import Optimizer
from ..tensor import Tensor
from .. import backend_api, cuda_graph


class SGD(Optimizer):
    def __init__(self, params: List[Tensor], lr: float = 1e-2, momentum: float = 0.0, weight_decay: float = 0.0,
                 nesterov: bool = False) -> None:
        super().__init__(params)
        self.lr, self.momentum, self.weight_decay, self.nesterov = lr, momentum, weight_decay, nesterov
        self.v = [backend_api.zeros(p.shape, device=p.device) for p in self.params]

    def step(self):
        grad_scale = self._grad_scale(fused=True)
        active = self._active()
        if not active:
            return
        dev = active[0][1].device
        for i, p, _ in active:
            self.v[i] = self._state_like(self.v[i], p)
        dev.multi_sgd_step(
            [p.data._handle for _, p, _ in active], [(g._handle, g._offset) for _, _, g in active],
            [self.v[i]._handle for i, _, _ in active], [p.data.size for _, p, _ in active], float(self.lr),
            float(self.momentum), float(self.weight_decay), bool(self.nesterov), float(grad_scale))
        cuda_graph.note_optimizer_step(self)

    def _graph_refresh(self, dev, graph_exec, index):
        dev.graph_set_sgd(graph_exec, index, float(self.lr), float(self.momentum), float(self.weight_decay),
                          bool(self.nesterov), float(self._grad_scale_value()))
