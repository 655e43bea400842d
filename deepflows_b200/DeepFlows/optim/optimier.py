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
        from .. import dist
        dist.pre_step()  # data parallel: gradients may alias buckets whose all-reduce is still in flight
        for param in self.params:
            param.zero_grad()

    # ---- shared by the fused steps ----------------------------------------------------------------------
    def _active(self):
        """(index, param, grad) for every parameter that has a gradient. The fused kernels walk parameter,
        gradient and optimizer state as flat buffers side by side, so the parameter storage is made dense
        (once; any permutation of a compact array, e.g. channels-last conv weights, qualifies) and the
        gradient is brought to the parameter's memory layout."""
        out = []
        for i, p in enumerate(self.params):
            g = p.grad
            if g is None:
                continue
            if hasattr(g, "data") and not hasattr(g, "_handle"):
                g = g.data
            if not p.data.is_dense():
                p.data = p.data.compact()
            out.append((i, p, g.with_layout_of(p.data)))
        return out

    @staticmethod
    def _state_like(state, p):
        """Optimizer state in the parameter's current memory layout (it changes when a conv weight is re-laid
        out channels-last at its first use, or when a checkpoint is loaded)."""
        return state if state.strides == p.data.strides and state.is_dense() else state.with_layout_of(p.data)

    @staticmethod
    def _grad_scale(fused=False):
        from .. import dist
        return dist.pre_step(fused)

    @staticmethod
    def _grad_scale_value():
        """1/world_size without touching the streams (graph replays already contain the wait)."""
        from .. import dist
        return 1.0 / dist.get_world_size()
