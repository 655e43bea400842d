"""Parameter: a Tensor that always requires grad and is registered by Module.__setattr__
(reference: DeepFlows/nn/parameter.py:5-34)."""
from ..tensor import Tensor
from .. import backend_api


class Parameter(Tensor):
    def __init__(self, data: Tensor):
        super().__init__(array=data.data, dtype=data.dtype, device=data.device, requires_grad=True)
        # a Parameter is a leaf even when created under no_grad (e.g. inside a module built in eval mode)
        if not self.requires_grad:
            self.requires_grad = True
            from ..tensor import Graph
            Graph.add(self)

    def requires_grad_(self, requires_grad: bool = True):
        self.requires_grad = bool(requires_grad)
        return self

    def to(self, device):
        name = device if isinstance(device, str) else device.name
        if self.device.name == name:
            return self
        return self.__class__(Tensor(self.data.numpy(), dtype=self.dtype, device=backend_api.Device(name)))

    def __repr__(self) -> str:
        return "Parameter : \n{},\ndevice={}".format(self.data, self.device)
