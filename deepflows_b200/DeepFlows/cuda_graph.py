"""CUDA-graph replay of a whole training step (new: the reference launches every op from Python, one
blocking kernel at a time, DeepFlows/backend/backend_src/ndarray_backend_cuda.cu passim).

    step = CapturedStep(lambda: train_step(model, opt, crit, x, t))   # x, t: device-resident Tensors
    for batch in loader:
        x.data.copy_from_host(batch)      # refresh the captured input buffers (outside the graph)
        loss = step()                     # 1st call: eager warm-up, 2nd: capture + launch, then: replay

The callable runs unchanged DeepFlows code (forward, loss, `backward()`, data-parallel all-reduce,
`optimizer.step()`); while it is being captured nothing executes, the kernels are recorded with the
buffers they touched, and `libdfb200` keeps those buffers reserved for the graph. A replay is ONE
`cudaGraphLaunch`: no Python autograd, no per-kernel launch cost. What may change between replays:
  * the contents of tensors that existed before the capture (inputs, targets, parameters, optimizer state);
  * optimizer hyper-parameters (`optimizer.lr` set by a scheduler, Adam's step counter): optimizers that
    stepped during the capture are refreshed through `dfb_graph_set_adam` / `dfb_graph_set_sgd`.
  * host-drawn tensors registered with `note_host_refill` (Dropout masks): their buffers are refilled from the host
    before every launch, in registration order, so a seeded run draws the same masks as an eager run.
What may not: shapes, control flow, other host<->device copies or `.numpy()` inside the callable (they raise).
"""
from . import backend_api

_active = None  # the CapturedStep being captured, if any


def capturing():
    return _active is not None


def note_optimizer_step(optimizer):
    """Called by fused optimizers from `step()`: remembers the order of optimizer steps in a capture."""
    if _active is not None:
        _active._optimizers.append(optimizer)


def note_host_refill(buffer, draw):
    """Called from inside a capture: `buffer` (a BackendTensor allocated during the capture) must hold `draw()` (a
    float32 numpy array of its shape) before every launch of the captured graph."""
    assert _active is not None, "note_host_refill outside a capture"
    _active._refills.append((buffer, draw))


class CapturedStep:
    def __init__(self, fn, device=None, warmup=1):
        self.fn = fn
        self.device = device if device is not None else backend_api.cuda()
        self.warmup = int(warmup)
        self.calls = 0
        self.result = None
        self._exec = None
        self._optimizers = []
        self._refills = []

    def _refill(self):
        for buf, draw in self._refills:
            self.device.from_numpy(draw(), buf._handle)

    @property
    def captured(self):
        return self._exec is not None

    def __call__(self):
        global _active
        self.calls += 1
        if self._exec is None:
            if self.calls <= self.warmup:
                return self.fn()
            dev = self.device
            dev.graph_begin_capture()
            _active = self
            try:
                self.result = self.fn()
            except BaseException:
                _active = None
                try:
                    dev.graph_destroy(dev.graph_end_capture())
                except Exception:
                    pass
                raise
            _active = None
            self._exec = dev.graph_end_capture()
            self._refill()
            dev.graph_launch(self._exec)  # the capture itself executed nothing
            return self.result
        for i, opt in enumerate(self._optimizers):
            opt._graph_refresh(self.device, self._exec, i)
        self._refill()
        self.device.graph_launch(self._exec)
        return self.result

    def node_counts(self):
        """(kernel nodes, all nodes) replayed by one launch of the captured graph."""
        return tuple(self.device.graph_node_counts(self._exec)) if self._exec is not None else (0, 0)

    def destroy(self):
        if self._exec is not None:
            self.device.graph_destroy(self._exec)
            self._exec = None
            self.result = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass
