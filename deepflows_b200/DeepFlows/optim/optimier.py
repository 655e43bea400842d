"""Optimizer base class (reference: DeepFlows/optim/optimier.py:6-15; the module name keeps the
reference's spelling so `from DeepFlows.optim.optimier import Optimizer` still works)."""
from typing import List

from ..tensor import Tensor


class Optimizer:
    def __init__(self, params: List[Tensor]) -> None:
        self.params: List[Tensor] = list(params)

    def step(self):
        raise NotImplementedError

    def zero_grad(self):
        for param in self.params:
            param.zero_grad()

    # ---- shared by the fused steps ----------------------------------------------------------------------
    def _active(self):
        """(index, param, compact grad) for every parameter that has a gradient. Parameter storage is
        made compact (once) so the fused kernel can update it in place."""
        out = []
        for i, p in enumerate(self.params):
            g = p.grad
            if g is None:
                continue
            if hasattr(g, "data") and not hasattr(g, "_handle"):
                g = g.data
            if not p.data.is_compact():
                p.data = p.data.compact()
            out.append((i, p, g.compact()))
        return out

    @staticmethod
    def _grad_scale():
        from .. import dist
        return dist.pre_step()

    @staticmethod
    def _grad_scale_value():
        """1/world_size without touching the streams (graph replays already contain the wait)."""
        from .. import dist
        return 1.0 / dist.get_world_size()
