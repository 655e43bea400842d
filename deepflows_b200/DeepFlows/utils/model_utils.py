"""Checkpoint save / load with the reference's on-disk layout
(reference: DeepFlows/utils/model_utils.py:19-181): a pickled dict
  {epoch, loss, model_parameters{name -> ndarray},
   optimizer_state{type, lr, momentum, weight_decay, v[], s[], t}}
plus, as a superset, `model_buffers{name -> ndarray}` so BatchNorm running statistics survive a
round trip (they are lost in the reference). Files written by the reference load here
and vice versa (the extra key is ignored there)."""
import os
import pickle
from typing import Any, Dict, Optional

import numpy as np

from ..nn.modules.module import Module
from ..optim.optimier import Optimizer
from .. import backend_api


def _host(t):
    return (t.data if hasattr(t, "data") and not hasattr(t, "_handle") else t).numpy().copy()


def save_checkpoint(model: Module, optimizer: Optional[Optimizer] = None, epoch: int = 0,
                    loss: Optional[float] = None, save_path: str = "checkpoint.pkl") -> None:
    os.makedirs(os.path.dirname(os.path.abspath(save_path)), exist_ok=True)
    ckpt = {"epoch": epoch, "loss": loss,
            "model_parameters": {name: _host(p) for name, p in model.named_parameters()},
            "model_buffers": {name: _host(b) for name, b in model.named_buffers()}}
    if optimizer is not None:
        state = {"type": type(optimizer).__name__}
        for attr in ("lr", "momentum", "weight_decay"):
            if hasattr(optimizer, attr):
                state[attr] = getattr(optimizer, attr)
        for attr in ("v", "s"):
            if hasattr(optimizer, attr):
                state[attr] = [_host(x) for x in getattr(optimizer, attr)]
        if hasattr(optimizer, "t"):
            state["t"] = optimizer.t
        ckpt["optimizer_state"] = state
    with open(save_path, "wb") as f:
        pickle.dump(ckpt, f)


def load_checkpoint(model: Module, optimizer: Optional[Optimizer] = None,
                    save_path: str = "checkpoint.pkl") -> Dict[str, Any]:
    if not os.path.exists(save_path):
        raise FileNotFoundError("checkpoint file does not exist: {}".format(save_path))
    with open(save_path, "rb") as f:
        ckpt = pickle.load(f)
    saved = ckpt.get("model_parameters", {})
    for name, p in model.named_parameters():
        if name in saved and isinstance(saved[name], np.ndarray):
            p.data = backend_api.Btensor(saved[name], dtype="float32", device=p.device)
    saved = ckpt.get("model_buffers", {})
    for name, b in model.named_buffers():
        if name in saved and isinstance(saved[name], np.ndarray):
            b.data = backend_api.Btensor(saved[name], dtype="float32", device=b.device)
    state = ckpt.get("optimizer_state")
    if optimizer is not None and state is not None:
        for attr in ("lr", "momentum", "weight_decay"):
            if attr in state and hasattr(optimizer, attr):
                setattr(optimizer, attr, state[attr])
        for attr in ("v", "s"):
            if attr in state and hasattr(optimizer, attr):
                slots = getattr(optimizer, attr)
                for i, p in enumerate(optimizer.params):
                    if i < len(state[attr]) and isinstance(state[attr][i], np.ndarray):
                        slots[i] = backend_api.Btensor(state[attr][i], dtype="float32", device=p.device)
        if "t" in state and hasattr(optimizer, "t"):
            optimizer.t = state["t"]
    return {"epoch": ckpt.get("epoch", 0), "loss": ckpt.get("loss")}
