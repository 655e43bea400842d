"""Dropout (reference: DeepFlows/nn/modules/dropout.py:7-35). The mask comes from the host RNG
(`np.random.binomial`) and is uploaded each step, exactly like the reference, so seeded runs draw the
same masks. Evaluation multiplies by (1 - p) (reference quirk Q8)."""
import numpy as np

from .module import Module


class Dropout(Module):
    def __init__(self, p: float = 0.5):
        super().__init__()
        assert 0 <= p < 1
        self.p = p

    def forward(self, x):
        if not self.training:
            return x * (1 - self.p)
        mask = x.device.empty(x.shape, dtype="float32")
        from ... import cuda_graph
        if cuda_graph.capturing():
            # Inside a captured step nothing may be copied from the host: the mask buffer belongs to the graph and is
            # refilled from the host RNG (same call, same order as an eager step) before every launch of the graph.
            cuda_graph.note_host_refill(mask, lambda shape=x.shape, p=self.p: np.random.binomial(1, 1 - p, shape).astype(np.float32))
        else:
            host = np.random.binomial(1, 1 - self.p, x.shape).astype(np.float32)
            x.device.from_numpy(host, mask._handle)
        return x * mask / (1 - self.p)

    def __repr__(self):
        return "{}(p={})".format(type(self).__name__, self.p)
