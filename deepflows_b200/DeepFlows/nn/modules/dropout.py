"""Dropout (reference: DeepFlows/nn/modules/dropout.py:7-35). The mask comes from the host RNG
(`np.random.binomial`) and is uploaded each step, exactly like the reference, so seeded runs draw the
same masks. Evaluation multiplies by (1 - p) (reference quirk Q8).

Opt-in (DEEPFLOWS_DROPOUT=device / backend_api.set_dropout_rng("device")): the mask is drawn on the device by a Philox
kernel (dfb_dropout_mask) - numpy's binomial costs ~10 ns per element on one host core, i.e. 5 ms for the 256 x 2048
activation of the CNN-CIFAR10 script, ten times the rest of its training step. The module's seed comes from numpy's
generator at first use (a seeded run stays reproducible); the masks are NOT the reference's."""
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
        from ...backend.backend_tensor import get_dropout_rng
        if get_dropout_rng() == "device" and x.device.has("dropout_mask"):
            if getattr(self, "_rng_seed", None) is None:
                object.__setattr__(self, "_rng_seed", int(np.random.randint(0, 1 << 24)))
                object.__setattr__(self, "_rng_step", 0)
            state = x.device.empty((2,), dtype="float32")

            def draw(mod=self):
                object.__setattr__(mod, "_rng_step", (mod._rng_step + 1) % (1 << 24))
                return np.array([mod._rng_seed, mod._rng_step], dtype=np.float32)

            if cuda_graph.capturing():      # 8 bytes from the host before every replay instead of the whole mask
                cuda_graph.note_host_refill(state, draw)
            else:
                x.device.from_numpy(draw(), state._handle)
            x.device.dropout_mask(mask._handle, mask.size, 1.0 - self.p, state._handle)
            return x * mask / (1 - self.p)
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
