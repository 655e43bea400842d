"""Backend selection, same exported names as the reference (DeepFlows/backend_selection.py:4-18):
`backend_api` is the ndarray module, `Device` is the BackendDevice *class*, `BackendTensor` the array."""
BACKEND = "nd"

from . import backend as backend_api  # noqa: E402
from .backend import (all_devices, cuda, cpu, cpu_numpy, gpu_cupy, default_device,  # noqa: E402,F401
                      BackendDevice as Device)

BackendTensor = backend_api.BackendTensor
