"""Adam (reference: DeepFlows/optim/adam.py:8-63): L2 weight decay folded into the gradient, first /
second moments `v` / `s`, bias correction with the step counter `t` starting at 1, eps added outside
the square root. The reference issues ~14 out-of-place BackendTensor kernels per parameter; here all
parameters are updated in place by one `multi_adam_step` launch."""
from .optimier import Optimizer
from .. import backend_api, cuda_graph


class Adam(Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay=0) -> None:
        super().__init__(params)
        self.lr = lr
        self.beta1, self.beta2 = betas
        self.eps = eps
        self.weight_decay = weight_decay
        self.v = [backend_api.zeros_like(p.data) for p in self.params]
        self.s = [backend_api.zeros_like(p.data) for p in self.params]
        self.t = 1

    def step(self):
        grad_scale = self._grad_scale(fused=True)
        active = self._active()
        if active:
            dev = active[0][1].device
            for i, p, _ in active:
                self.v[i] = self._state_like(self.v[i], p)
                self.s[i] = self._state_like(self.s[i], p)
            dev.multi_adam_step(
                [p.data._handle for _, p, _ in active], [(g._handle, g._offset) for _, _, g in active],
                [self.v[i]._handle for i, _, _ in active], [self.s[i]._handle for i, _, _ in active],
                [p.data.size for _, p, _ in active], float(self.lr), float(self.beta1), float(self.beta2),
                float(self.eps), float(self.weight_decay), int(self.t), float(grad_scale))
            for _, p, _ in active:
                p.children.clear()
                p.parents.clear()
            cuda_graph.note_optimizer_step(self)
        self.t += 1

    def _graph_refresh(self, dev, graph_exec, index):
        """Hyper-parameters for the next replay of a captured step (cuda_graph.CapturedStep)."""
        dev.graph_set_adam(graph_exec, index, float(self.lr), float(self.beta1), float(self.beta2), float(self.eps),
                           float(self.weight_decay), int(self.t), float(self._grad_scale_value()))
        self.t += 1
