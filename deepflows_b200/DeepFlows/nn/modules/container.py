"""Containers (reference: DeepFlows/nn/modules/container.py:10-112; ModuleList / ModuleDict are empty
stubs there and real containers here)."""
from collections import OrderedDict
from typing import Iterator

from .module import Module


class Sequential(Module):
    def __init__(self, *args):
        super().__init__()
        if len(args) == 1 and isinstance(args[0], OrderedDict):
            for key, module in args[0].items():
                self.add_module(key, module)
        else:
            for idx, module in enumerate(args):
                self.add_module(str(idx), module)

    def __len__(self):
        return len(self._modules)

    def __getitem__(self, idx):
        if isinstance(idx, slice):
            return Sequential(OrderedDict(list(self._modules.items())[idx]))
        return list(self._modules.values())[idx]

    def __add__(self, other):
        if not isinstance(other, Sequential):
            raise ValueError("add operator supports only objects of Sequential class, but {} is given.".format(type(other)))
        out = Sequential()
        for layer in list(self) + list(other):
            out.append(layer)
        return out

    def __iter__(self) -> Iterator[Module]:
        return iter(self._modules.values())

    def forward(self, input):
        for module in self:
            input = module(input)
        return input

    def append(self, module: Module) -> "Sequential":
        self.add_module(str(len(self)), module)
        return self

    def extend(self, sequential) -> "Sequential":
        for layer in sequential:
            self.append(layer)
        return self


class ModuleList(Module):
    def __init__(self, modules=None):
        super().__init__()
        for m in modules or []:
            self.append(m)

    def append(self, module: Module):
        self.add_module(str(len(self._modules)), module)
        return self

    def extend(self, modules):
        for m in modules:
            self.append(m)
        return self

    def __len__(self):
        return len(self._modules)

    def __getitem__(self, idx):
        return list(self._modules.values())[idx]

    def __iter__(self):
        return iter(self._modules.values())


class ModuleDict(Module):
    def __init__(self, modules=None):
        super().__init__()
        for k, m in (modules or {}).items():
            self.add_module(k, m)

    def __getitem__(self, key):
        return self._modules[key]

    def __setitem__(self, key, module):
        self.add_module(key, module)

    def __len__(self):
        return len(self._modules)

    def __iter__(self):
        return iter(self._modules)

    def keys(self):
        return self._modules.keys()

    def items(self):
        return self._modules.items()

    def values(self):
        return self._modules.values()
