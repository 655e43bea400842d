"""nn.Module: parameter / buffer / sub-module registries with the torch-like surface the reference
exposes (DeepFlows/nn/modules/module.py:32-856).

Behaviours kept on purpose: attribute assignment registers only `Parameter` and `Module` values, so
modules held in plain Python lists stay invisible to `parameters()` (reference lines
373-444); `train()` / `eval()` also flip the process-global autograd switch (Q9, line 764).
"""
from collections import OrderedDict
from copy import deepcopy
from typing import Any, Callable, Dict, Iterator, Optional, Set, Tuple

import numpy as np

from ...autograd import set_grad_enabled
from ...tensor import Tensor
from ... import backend_api
from ..parameter import Parameter


def _indent(text, n):
    lines = text.split("\n")
    if len(lines) == 1:
        return text
    return lines[0] + "\n" + "\n".join(" " * n + ln for ln in lines[1:])


class Module:
    training: bool

    def __init__(self) -> None:
        d = self.__dict__
        d["training"] = True
        d["_parameters"] = OrderedDict()
        d["_buffers"] = OrderedDict()
        d["_non_persistent_buffers_set"] = set()
        d["_modules"] = OrderedDict()

    def forward(self, *input):
        raise NotImplementedError('Module [{}] is missing the required "forward" function'.format(type(self).__name__))

    def __call__(self, *input):
        return self.forward(*input)

    # ---- registration -------------------------------------------------------------------------------
    def _check_name(self, kind, name, own_registry):
        if not isinstance(name, str):
            raise TypeError("{} name should be a string. Got {}".format(kind, type(name).__name__))
        if "." in name:
            raise KeyError('{} name can\'t contain "."'.format(kind))
        if name == "":
            raise KeyError('{} name can\'t be empty string ""'.format(kind))
        if hasattr(self, name) and name not in own_registry:
            raise KeyError("attribute '{}' already exists".format(name))

    def register_buffer(self, name: str, tensor: Optional[Tensor], persistent: bool = True) -> None:
        if "_buffers" not in self.__dict__:
            raise AttributeError("cannot assign buffer before Module.__init__() call")
        self._check_name("buffer", name, self._buffers)
        if tensor is not None and not isinstance(tensor, Tensor):
            raise TypeError("cannot assign '{}' object to buffer '{}' (Tensor or None required)".format(
                type(tensor).__name__, name))
        self._buffers[name] = tensor
        if persistent:
            self._non_persistent_buffers_set.discard(name)
        else:
            self._non_persistent_buffers_set.add(name)

    def register_parameter(self, name: str, param: Optional[Parameter]) -> None:
        if "_parameters" not in self.__dict__:
            raise AttributeError("cannot assign parameter before Module.__init__() call")
        self._check_name("parameter", name, self._parameters)
        if param is not None and not isinstance(param, Parameter):
            raise TypeError("cannot assign '{}' object to parameter '{}' (nn.Parameter or None required)".format(
                type(param).__name__, name))
        self._parameters[name] = param

    def add_module(self, name: str, module: Optional["Module"]) -> None:
        if module is not None and not isinstance(module, Module):
            raise TypeError("{} is not a Module subclass".format(type(module).__name__))
        self._check_name("module", name, self._modules)
        self._modules[name] = module

    register_module = add_module

    def __getattr__(self, name: str) -> Any:
        d = self.__dict__
        for registry in ("_parameters", "_buffers", "_modules"):
            reg = d.get(registry)
            if reg is not None and name in reg:
                return reg[name]
        raise AttributeError("'{}' object has no attribute '{}'".format(type(self).__name__, name))

    def __setattr__(self, name: str, value) -> None:
        d = self.__dict__
        params, buffers, modules = d.get("_parameters"), d.get("_buffers"), d.get("_modules")

        def forget(*registries):
            for reg in registries:
                if reg is not None and name in reg:
                    if isinstance(reg, set):
                        reg.discard(name)
                    else:
                        del reg[name]

        if isinstance(value, Parameter):
            if params is None:
                raise AttributeError("cannot assign parameters before Module.__init__() call")
            forget(d, buffers, modules, d.get("_non_persistent_buffers_set"))
            self.register_parameter(name, value)
        elif params is not None and name in params:
            if value is not None:
                raise TypeError("cannot assign this value as parameter '{}' (Parameter or None expected)".format(name))
            self.register_parameter(name, value)
        elif isinstance(value, Module):
            if modules is None:
                raise AttributeError("cannot assign module before Module.__init__() call")
            forget(d, params, buffers, d.get("_non_persistent_buffers_set"))
            modules[name] = value
        elif modules is not None and name in modules:
            if value is not None:
                raise TypeError("cannot assign this as child module '{}' (nn.Module or None expected)".format(name))
            modules[name] = value
        elif buffers is not None and name in buffers:
            if value is not None and not isinstance(value, Tensor):
                raise TypeError("cannot assign this as buffer '{}' (Tensor or None expected)".format(name))
            buffers[name] = value
        else:
            object.__setattr__(self, name, value)

    def __delattr__(self, name):
        if name in self._parameters:
            del self._parameters[name]
        elif name in self._buffers:
            del self._buffers[name]
            self._non_persistent_buffers_set.discard(name)
        elif name in self._modules:
            del self._modules[name]
        else:
            object.__delattr__(self, name)

    # ---- lookup by dotted path ------------------------------------------------------------------------
    def get_submodule(self, target: str) -> "Module":
        mod = self
        if target == "":
            return mod
        for item in target.split("."):
            if not hasattr(mod, item):
                raise AttributeError(mod._get_name() + " has no attribute `" + item + "`")
            mod = getattr(mod, item)
            if not isinstance(mod, Module):
                raise AttributeError("`" + item + "` is not an nn.Module")
        return mod

    def get_parameter(self, target: str) -> "Parameter":
        module_path, _, name = target.rpartition(".")
        mod = self.get_submodule(module_path)
        if not hasattr(mod, name):
            raise AttributeError(mod._get_name() + " has no attribute `" + name + "`")
        param = getattr(mod, name)
        if not isinstance(param, Parameter):
            raise AttributeError("`" + name + "` is not an nn.Parameter")
        return param

    def get_buffer(self, target: str) -> "Tensor":
        module_path, _, name = target.rpartition(".")
        mod = self.get_submodule(module_path)
        if name not in mod._buffers:
            raise AttributeError("`" + name + "` is not a buffer")
        return mod._buffers[name]

    def apply(self, fn: Callable[["Module"], None]):
        for module in self.children():
            module.apply(fn)
        fn(self)
        return self

    # ---- iteration ------------------------------------------------------------------------------------
    def named_modules(self, memo: Optional[Set["Module"]] = None, prefix: str = "", remove_duplicate: bool = True):
        if memo is None:
            memo = set()
        if self not in memo:
            if remove_duplicate:
                memo.add(self)
            yield prefix, self
            for name, module in self._modules.items():
                if module is None:
                    continue
                yield from module.named_modules(memo, prefix + ("." if prefix else "") + name, remove_duplicate)

    def modules(self) -> Iterator["Module"]:
        for _, module in self.named_modules():
            yield module

    def named_children(self) -> Iterator[Tuple[str, "Module"]]:
        seen = set()
        for name, module in self._modules.items():
            if module is not None and module not in seen:
                seen.add(module)
                yield name, module

    def children(self) -> Iterator["Module"]:
        for _, module in self.named_children():
            yield module

    def _named_members(self, get_members_fn, prefix="", recurse=True, remove_duplicate: bool = True):
        seen = set()
        mods = self.named_modules(prefix=prefix, remove_duplicate=remove_duplicate) if recurse else [(prefix, self)]
        for mod_prefix, mod in mods:
            for k, v in get_members_fn(mod):
                if v is None or id(v) in seen:
                    continue
                if remove_duplicate:
                    seen.add(id(v))
                yield mod_prefix + ("." if mod_prefix else "") + k, v

    def named_parameters(self, prefix: str = "", recurse: bool = True, remove_duplicate: bool = True):
        yield from self._named_members(lambda m: m._parameters.items(), prefix, recurse, remove_duplicate)

    def parameters(self, recurse: bool = True) -> Iterator[Parameter]:
        for _, p in self.named_parameters(recurse=recurse):
            yield p

    def named_buffers(self, prefix: str = "", recurse: bool = True, remove_duplicate: bool = True):
        yield from self._named_members(lambda m: m._buffers.items(), prefix, recurse, remove_duplicate)

    def buffers(self, recurse: bool = True) -> Iterator[Tensor]:
        for _, b in self.named_buffers(recurse=recurse):
            yield b

    # ---- state ----------------------------------------------------------------------------------------
    def params_and_buffers_saved(self):
        out = {k: v for k, v in self._parameters.items() if v is not None}
        out.update({k: v for k, v in self._buffers.items()
                    if v is not None and k not in self._non_persistent_buffers_set})
        return out

    def state_dict(self) -> Dict[str, np.ndarray]:
        """name -> host ndarray for every parameter and persistent buffer (new; the reference has only
        `named_parameters()` + utils.model_utils pickling)."""
        out = OrderedDict()
        for prefix, mod in self.named_modules():
            for k, v in mod.params_and_buffers_saved().items():
                out[prefix + ("." if prefix else "") + k] = v.data.numpy().copy()
        return out

    def load_state_dict(self, state_dict: Dict[str, Any], strict: bool = True):
        """Reference: nn/modules/module.py:471-537 (same assignment rules and error text). One deliberate
        difference: BatchNorm running statistics are registered buffers here (so they are saved and loaded),
        plain attributes in the reference; a dict without them - anything the reference wrote - therefore
        still loads under strict=True, i.e. missing *buffers* are not an error."""
        missing, unexpected = [], set(state_dict.keys())

        def assign(target, value):
            if isinstance(value, Tensor):
                target.data = value.data
            else:
                host = value.numpy() if hasattr(value, "numpy") and not isinstance(value, np.ndarray) else value
                target.data = backend_api.Btensor(np.asarray(host), device=target.device, dtype=target.dtype)

        for prefix, mod in self.named_modules():
            for registry in (mod._parameters, mod._buffers):
                for name, t in registry.items():
                    if t is None:
                        continue
                    key = prefix + ("." if prefix else "") + name
                    if key in state_dict:
                        assign(t, state_dict[key])
                        unexpected.discard(key)
                    elif registry is mod._parameters:
                        missing.append(key)
        if strict:
            errors = []
            if unexpected:
                errors.append("Unexpected key(s) in state_dict: {}.".format(", ".join(sorted(unexpected))))
            if missing:
                errors.append("Missing key(s) in state_dict: {}.".format(", ".join(missing)))
            if errors:
                raise RuntimeError("Error(s) in loading state_dict for {}:\n\t{}".format(
                    type(self).__name__, "\n\t".join(errors)))

    def load_weights(self, weights: Dict[str, Any]):
        self.load_state_dict(weights, strict=False)

    # ---- modes ----------------------------------------------------------------------------------------
    def train(self, mode: bool = True):
        if not isinstance(mode, bool):
            raise ValueError("training mode is expected to be boolean")
        self.training = mode
        set_grad_enabled(mode)  # reference quirk Q9
        for module in self.children():
            module.train(mode)
        return self

    def eval(self):
        return self.train(False)

    def requires_grad_(self, requires_grad: bool = True):
        for p in self.parameters():
            p.requires_grad_(requires_grad)
        return self

    def zero_grad(self, set_to_none: bool = True) -> None:
        for p in self.parameters():
            if p.grad is not None:
                if set_to_none:
                    p.grad = None
                else:
                    p.grad.fill(0.0)

    # ---- device moves ---------------------------------------------------------------------------------
    def to(self, device):
        """A copy of this module on `device` (self when every parameter and buffer already lives there); the
        reference deep-copies and re-homes the copy [806-818]. Tensors are rebuilt through numpy on the target
        device, parameters stay Parameters, registered buffers (BatchNorm running statistics) move too."""
        name = device if isinstance(device, str) else device.name
        tensors = [t for t in list(self.parameters()) + list(self.buffers()) if t is not None]
        if all(t.device.name == name for t in tensors):
            object.__setattr__(self, "device", backend_api.Device(name))
            return self
        module = deepcopy(self)
        module.move(name)
        return module

    def move(self, device):
        name = device if isinstance(device, str) else device.name
        object.__setattr__(self, "device", backend_api.Device(name))
        for key, p in list(self._parameters.items()):
            if p is not None:
                self._parameters[key] = p.to(name)
        for key, b in list(self._buffers.items()):
            if b is not None:
                self._buffers[key] = b.to(name)
        for module in self._modules.values():
            if isinstance(module, Module):
                module.move(name)

    def cuda(self):
        return self.to("cuda")

    def cpu(self):
        return self.to("cpu")

    # ---- printing -------------------------------------------------------------------------------------
    def _get_name(self):
        return type(self).__name__

    def extra_repr(self) -> str:
        return ""

    def __repr__(self):
        head = self._get_name() + "("
        extra = self.extra_repr()
        child_lines = ["(" + k + "): " + _indent(repr(m), 2) for k, m in self._modules.items()]
        if not child_lines:
            return head + extra + ")"
        body = ([extra] if extra else []) + child_lines
        return head + "\n  " + "\n  ".join(body) + "\n)"
