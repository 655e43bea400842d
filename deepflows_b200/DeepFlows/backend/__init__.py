from .backend_tensor import *  # noqa: F401,F403
from . import backend_tensor as _bt

# names that `import *` skips but callers reach through `backend_api.<name>`
from .backend_tensor import (BackendDevice, BackendTensor, Device, all_devices, cuda, cpu, cpu_numpy,
                             gpu_cupy, default_device, register_numpy_device, set_precision, get_precision,
                             set_dgrad_mode, get_dgrad_mode)
