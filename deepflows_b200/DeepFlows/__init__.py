"""DeepFlows host package for the B200 backend (deepflows_b200).

Same import surface as the reference package (DeepFlows/__init__.py:1-3): `from DeepFlows import
tensor, nn, backend_api, ...`. The only compute device is `cuda` = libdfb200.so through the
`CUDA_BACKEND` shim; there is no CPU compute path in this package."""
from .tensor import *  # noqa: F401,F403
from .autograd import enable_grad, no_grad  # noqa: F401
from .backend_selection import *  # noqa: F401,F403
from .backend_selection import backend_api, BackendTensor, Device  # noqa: F401
