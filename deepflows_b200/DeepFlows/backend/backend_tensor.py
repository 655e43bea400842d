"""Strided-view ndarray over opaque float32 device buffers, and the device registry.

Host-side mirror of the reference's `backend_tensor.py` (DeepFlows/backend/backend_tensor.py):
same class and function names, same view algebra (shape / strides / offset in elements over a
flat handle), same device-module protocol underneath (`device.<op>(handle, ..., out_handle)`,
28 flat functions, reference lines 64-172 for the protocol and ndarray_backend_cuda.cu:515-716 for
the CUDA module). Differences that matter for speed, none for results:

* ops on *dense* operands with identical strides run on the raw buffers without compacting, and
  the result keeps those strides (channels-last activations stay channels-last);
* `x + bias` with a per-channel / per-column bias uses one fused kernel instead of
  broadcast_to + compact + ewise_add (reference lines 533-542);
* the only device shipped is `cuda` (libdfb200.so). `cpu` exists only when a test registers a
  numpy device module with `register_numpy_device` - this package never computes on the host.
"""
import operator
import os
from functools import reduce

import numpy as np

__all__ = [
    "prod", "BackendDevice", "BackendTensor", "cuda", "cpu", "cpu_numpy", "gpu_cupy", "default_device",
    "all_devices", "Device", "Btensor", "empty", "full", "zeros", "ones", "zeros_like", "ones_like",
    "broadcast_to", "reshape", "maximum", "max", "log", "exp", "tanh", "flip", "summation", "mean", "pad",
    "expand_dims", "register_numpy_device", "set_precision", "get_precision", "set_dgrad_mode",
    "get_dgrad_mode", "set_fix_mean", "get_fix_mean", "set_fusion", "get_fusion", "PendingTensor", "BnApply",
    "set_dropout_rng", "get_dropout_rng",
]


def prod(x):
    return reduce(operator.mul, x, 1)


# ------------------------------------------------------------------------------------------------
# devices
# ------------------------------------------------------------------------------------------------
class BackendDevice:
    """A named device wrapping the module that implements the op protocol (reference lines 11-51)."""

    def __init__(self, name, mod):
        self.name = name
        self.mod = mod

    def __eq__(self, other):
        return isinstance(other, BackendDevice) and self.name == other.name

    def __hash__(self):
        return hash(self.name)

    def __repr__(self):
        return self.name + "()"

    def __getattr__(self, attr):
        # only reached for names that are not instance attributes: forward to the module once and
        # remember the bound function so later lookups are plain attribute hits
        if attr.startswith("__"):
            raise AttributeError(attr)
        mod = self.__dict__.get("mod")
        if mod is None:
            raise RuntimeError(
                "device '%s' is not available: %s" % (self.__dict__.get("name"), _unavailable_reason(self.__dict__.get("name"))))
        fn = getattr(mod, attr)
        self.__dict__[attr] = fn
        return fn

    def enabled(self):
        return self.mod is not None

    def __deepcopy__(self, memo):
        return self  # a device is a process-wide singleton (its module object cannot be copied)

    def __reduce__(self):
        return (Device, (self.name,))

    def has(self, attr):
        """True when the device module implements the (fused, L1) entry point `attr` (remembered: the host layer asks a
        couple of hundred times per training step, and a module's entry points do not change)."""
        d = self.__dict__
        cache = d.get("_has_cache")
        if cache is None:
            cache = d["_has_cache"] = {}
        v = cache.get(attr)
        if v is None:
            v = cache[attr] = self.mod is not None and hasattr(self.mod, attr)
        return v

    def randn(self, *shape, dtype="float32"):
        return BackendTensor(np.random.randn(*shape).astype(dtype), device=self)

    def rand(self, *shape, dtype="float32"):
        return BackendTensor(np.random.rand(*shape).astype(dtype), device=self)

    def one_hot(self, n, i, dtype="float32"):
        return BackendTensor(np.eye(n, dtype=dtype)[i], device=self)

    def empty(self, shape, dtype="float32"):
        assert dtype in (None, "float32")
        return BackendTensor.make(shape, device=self)

    def full(self, shape, fill_value, dtype="float32"):
        assert dtype in (None, "float32")
        out = BackendTensor.make(shape, device=self)
        out.fill(fill_value)
        return out


_cuda_device = None
_cuda_error = None
_numpy_device = None


def _unavailable_reason(name):
    if name == "cuda":
        return ("the CUDA_BACKEND extension could not be imported (%s). Build it with "
                "`python -m deepflows_b200.build`; there is no CPU fallback." % (_cuda_error,))
    return "no numpy device module is registered (the numpy device lives in oracle/ and is for tests only)"


def cuda():
    """The B200 device. Like the reference (lines 54-61) an import failure yields a disabled device
    object; unlike the reference, using it raises a RuntimeError that says why."""
    global _cuda_device, _cuda_error
    if _cuda_device is None:
        try:
            from DeepFlows.backend.backend_src.build.Release import CUDA_BACKEND
            _cuda_device = BackendDevice("cuda", CUDA_BACKEND)
        except ImportError as e:  # pragma: no cover - exercised only without a build
            _cuda_error = e
            return BackendDevice("cuda", None)
    return _cuda_device


def register_numpy_device(mod):
    """Test hook: install a module object implementing the device protocol on numpy arrays
    (oracle/numpy_device.py) as the `cpu` device. Pass None to remove it."""
    global _numpy_device
    _numpy_device = BackendDevice("cpu", mod) if mod is not None else None
    return _numpy_device


def cpu_numpy():
    return _numpy_device if _numpy_device is not None else BackendDevice("cpu", None)


def cpu():
    return cpu_numpy()


def gpu_cupy():
    return None


def default_device():
    """The reference defaults to its numpy device (line 185); this package defaults to the registered
    numpy device when a test installed one, else to cuda."""
    return _numpy_device if _numpy_device is not None else cuda()


def all_devices():
    return {"cpu": cpu(), "cuda": cuda(), "cpu_numpy": cpu_numpy(), "gpu_cupy": gpu_cupy()}


def Device(device_name=None):
    if isinstance(device_name, BackendDevice):
        return device_name
    if device_name == "cuda":
        return cuda()
    if device_name in ("cpu", "cpu_numpy"):
        return cpu_numpy()
    return all_devices()[device_name]


# ---- numerics switches (new; the reference has a single fp32 path) ---------------------------------
_PRECISIONS = {"fp32": 0, "tf32": 1, "bf16": 2, "simt": 3}
_precision = os.environ.get("DEEPFLOWS_PRECISION", "fp32").lower()
# 'exact' (the true transposed convolution) is the default; 'reference' reproduces the upstream last-writer-wins
# im2col backward bit for bit and is what the golden-parity tests select explicitly.
_dgrad_mode = os.environ.get("DEEPFLOWS_DGRAD", "exact").lower()


def set_precision(name):
    """Operand precision of matmul/conv: 'fp32' (exact FFMA), 'tf32' or 'bf16' (tcgen05, fp32 accumulate)."""
    global _precision
    name = name.lower()
    if name not in _PRECISIONS:
        raise ValueError("precision must be one of %s" % sorted(_PRECISIONS))
    _precision = name
    dev = cuda()
    if dev.enabled():
        dev.set_matmul_mode(_PRECISIONS[name])


def get_precision():
    return _precision


def precision_mode():
    return _PRECISIONS[_precision]


def set_dgrad_mode(name):
    """'exact' (default): the true transposed convolution. 'reference': the last-writer-wins input gradient of
    the reference's im2col backward (DeepFlows/nn/functional.py:285-294) - mathematically wrong for
    overlapping windows, kept as an opt-in bug-parity mode (also DEEPFLOWS_DGRAD=reference)."""
    global _dgrad_mode
    name = name.lower()
    if name not in ("reference", "exact"):
        raise ValueError("dgrad mode must be 'reference' or 'exact'")
    _dgrad_mode = name


def get_dgrad_mode():
    return _dgrad_mode


# `mean(axis)` of the reference divides by the TOTAL element count (backend_tensor.py:659-662), so the
# scripts' two-step global average pool returns true_mean / (N^2 C^2 W). That is kept by default (script parity);
# DEEPFLOWS_FIX_MEAN=1 / set_fix_mean(True) divides by the length of the reduced axis instead.
_fix_mean = os.environ.get("DEEPFLOWS_FIX_MEAN", "0") == "1"


def set_fix_mean(enabled):
    global _fix_mean
    _fix_mean = bool(enabled)


def get_fix_mean():
    return _fix_mean


# Cross-op fusion on the cuda device (new): convolutions hand the per-channel statistics of their output to the
# BatchNorm that follows, BatchNorm produces its output lazily so that `bn(x) + identity`, `bn(x) + bn_ds(x2)` and
# `relu(...)` of it become ONE kernel, conv dgrad adds the identity branch's gradient in its epilogue and computes the
# BatchNorm-backward reductions of the gradient it writes. DEEPFLOWS_FUSE=0 / set_fusion(False) runs every op on its own
# (what round 1 did; same results to rounding).
_fusion = os.environ.get("DEEPFLOWS_FUSE", "1") != "0"


def set_fusion(enabled):
    global _fusion
    _fusion = bool(enabled)


def get_fusion():
    return _fusion


# Where Dropout draws its masks: "host" = numpy's generator like the reference (seeded runs see the reference's masks),
# "device" = a Philox kernel (nn/modules/dropout.py)
_dropout_rng = os.environ.get("DEEPFLOWS_DROPOUT", "host").lower()


def set_dropout_rng(where):
    global _dropout_rng
    if where not in ("host", "device"):
        raise ValueError("dropout rng must be 'host' or 'device'")
    _dropout_rng = where


def get_dropout_rng():
    return _dropout_rng


# ------------------------------------------------------------------------------------------------
# BackendTensor
# ------------------------------------------------------------------------------------------------
def _compact_strides(shape):
    strides = [1] * len(shape)
    for i in range(len(shape) - 2, -1, -1):
        strides[i] = strides[i + 1] * shape[i + 1]
    return tuple(strides)


class BackendTensor:
    """N-d float32 array = (shape, strides, offset) view over a flat device handle."""

    # `_aux`: what a producer attached for the next consumer (fusion): ("colstats", mean_var Array) on a convolution's
    # output, ("bn_sums", sums Array, {id(record): row}) on a gradient whose BatchNorm-backward reductions exist
    __slots__ = ("_shape", "_strides", "_offset", "_device", "_handle", "_dense", "_aux")

    def __init__(self, other, device=None):
        if isinstance(other, BackendTensor):
            src = other.to(device if device is not None else other.device) + 0.0
        elif isinstance(other, np.ndarray):
            device = device if device is not None else default_device()
            src = BackendTensor.make(other.shape, device=device)
            host = np.ascontiguousarray(other, dtype=np.float32)
            src._device.from_numpy(host, src._handle)
        else:
            src = BackendTensor(np.array(other), device)
        self._adopt(src)

    def _adopt(self, src):
        self._shape = src._shape
        self._strides = src._strides
        self._offset = src._offset
        self._device = src._device
        self._handle = src._handle
        self._dense = src._dense
        self._aux = None

    # reference name for the same thing
    _init = _adopt

    @staticmethod
    def compact_strides(shape):
        return _compact_strides(shape)

    @staticmethod
    def make(shape, strides=None, device=None, handle=None, offset=0):
        """New view; allocates prod(shape) floats when no handle is given (reference lines 237-252)."""
        t = BackendTensor.__new__(BackendTensor)
        t._shape = tuple(int(s) for s in shape)
        t._strides = _compact_strides(t._shape) if strides is None else tuple(int(s) for s in strides)
        t._offset = int(offset)
        t._device = device if device is not None else default_device()
        t._handle = t._device.Array(prod(t._shape)) if handle is None else handle
        t._dense = None
        t._aux = None
        return t

    def __deepcopy__(self, memo):
        """Same values, same layout, a buffer of its own on the same device (used by Module.to / copy.deepcopy)."""
        src = self if self.is_dense() else self.compact()
        out = src._like()
        src._device.scalar_add(src._handle, 0.0, out._handle)
        return out

    # ---- properties ---------------------------------------------------------------------------------
    @property
    def shape(self):
        return self._shape

    @property
    def strides(self):
        return self._strides

    @property
    def device(self):
        return self._device

    @property
    def dtype(self):
        return "float32"

    @property
    def ndim(self):
        return len(self._shape)

    @property
    def size(self):
        return prod(self._shape)

    def __repr__(self):
        return "BackendTensor(" + str(self.numpy()) + ", device=%s)" % (self._device,)

    def __str__(self):
        return str(self.numpy())

    def __len__(self):
        return self._shape[0]

    # ---- layout predicates ----------------------------------------------------------------------------
    def is_compact(self):
        return self._strides == _compact_strides(self._shape) and prod(self._shape) == self._handle.size

    def is_dense(self):
        """True when the view is a permutation of a compact array that covers the whole handle, i.e.
        elementwise kernels may run on the raw buffer and keep the strides."""
        d = self._dense
        if d is None:
            d = False
            if self._offset == 0 and prod(self._shape) == self._handle.size:
                expect = 1
                d = True
                for st, sh in sorted(((st, sh) for st, sh in zip(self._strides, self._shape) if sh != 1)):
                    if st != expect:
                        d = False
                        break
                    expect *= sh
            self._dense = d
        return d

    def _dense_within_view(self):
        """The strides are a permutation of a compact layout of this shape (offset and handle size not considered):
        the elements occupy one contiguous run of prod(shape) floats starting at `_offset`."""
        expect = 1
        for st, sh in sorted(((st, sh) for st, sh in zip(self._strides, self._shape) if sh != 1)):
            if st != expect:
                return False
            expect *= sh
        return True

    def is_channels_last(self):
        """4-d (N,C,H,W) view whose memory order is N,H,W,C and which covers its handle."""
        if len(self._shape) != 4 or self._offset != 0:
            return False
        n, c, h, w = self._shape
        if n * c * h * w != self._handle.size:
            return False
        want = (h * w * c, 1, w * c, c)
        return all(sh == 1 or st == ws for st, ws, sh in zip(self._strides, want, self._shape))

    def channels_last(self):
        """This tensor (4-d, logical NCHW) with NHWC memory order; copies only when needed."""
        if self.is_channels_last():
            return self
        n, c, h, w = self._shape
        nhwc = self.permute((0, 2, 3, 1)).compact()
        return BackendTensor.make((n, c, h, w), (h * w * c, 1, w * c, c), self._device, nhwc._handle)

    def with_layout_of(self, ref):
        """This tensor's values in `ref`'s memory layout (same shape): returned as is when the strides already
        agree, otherwise copied into a new buffer with ref's strides. `ref` must be dense. Used wherever two
        tensors are walked as flat buffers side by side (fused optimizer steps, gradient buckets)."""
        assert self._shape == ref._shape, (self._shape, ref._shape)
        if self._strides == ref._strides and (self.is_dense() or self._dense_within_view()):
            return self  # may start at an offset inside a larger buffer (a gradient that aliases its all-reduce bucket)
        if ref.is_compact():
            return self.compact()
        if ref.is_channels_last():
            return self.channels_last()
        out = ref._like()
        out[tuple(slice(None) for _ in self._shape)] = self
        return out

    def flat_storage(self):
        """1-d view of the whole underlying buffer in memory order (the tensor must be dense)."""
        assert self.is_dense(), "flat_storage needs a dense tensor"
        return BackendTensor.make((self._handle.size,), None, self._device, self._handle, 0)

    # ---- basic manipulation ---------------------------------------------------------------------------
    def fill(self, value):
        self._device.fill(self._handle, value)

    def to(self, device):
        if device == self._device:
            return self
        return BackendTensor(self.numpy(), device=device)

    def numpy(self):
        return self._device.to_numpy(self._handle, self._shape, self._strides, self._offset)

    def compact(self):
        if self.is_compact():
            return self
        out = BackendTensor.make(self._shape, device=self._device)
        self._device.compact(self._handle, out._handle, self._shape, self._strides, self._offset)
        return out

    def as_strided(self, shape, strides):
        assert len(shape) == len(strides)
        return BackendTensor.make(shape, strides=strides, device=self._device, handle=self._handle)

    @property
    def flat(self):
        return self.reshape((self.size,))

    def reshape(self, new_shape):
        """View with a new shape over the same memory; requires a compact array. One -1 is allowed in
        the first or last position (reference lines 351-364)."""
        new_shape = tuple(new_shape)
        total = prod(self._shape)
        if new_shape and new_shape[0] == -1:
            rest = prod(new_shape[1:])
            new_shape = (int(total / rest),) + new_shape[1:]
        elif new_shape and new_shape[-1] == -1:
            rest = prod(new_shape[:-1])
            new_shape = new_shape[:-1] + (int(total / rest),)
        if prod(new_shape) != total:
            raise ValueError("Product of current shape is not equal to the product of the new shape!")
        if not self.is_compact():
            raise ValueError("The matrix is not compact!")
        return BackendTensor.make(new_shape, _compact_strides(new_shape), self._device, self._handle)

    def transpose(self, new_axes=None):
        if new_axes:
            axes = [new_axes[i] for i in range(self.ndim)]
        else:
            axes = list(range(self.ndim - 1, -1, -1))
        return self.permute(axes)

    def permute(self, new_axes):
        axes = [int(a) for a in new_axes]
        return BackendTensor.make([self._shape[a] for a in axes], [self._strides[a] for a in axes],
                                  self._device, self._handle, self._offset)

    def broadcast_to(self, new_shape):
        new_shape = tuple(new_shape)
        lead = len(new_shape) - len(self._shape)
        assert lead >= 0, "Cannot broadcast %s to %s" % (self._shape, new_shape)
        strides = [0] * lead
        for have, want, st in zip(self._shape, new_shape[lead:], self._strides):
            assert have == want or have == 1, "Dimension mismatch: %d vs %d" % (have, want)
            strides.append(st if have != 1 else 0)
        # a broadcast size-1 axis keeps stride 0 even when want == 1 (harmless: the index is always 0)
        return BackendTensor.make(new_shape, tuple(strides), self._device, self._handle, self._offset)

    # ---- indexing -------------------------------------------------------------------------------------
    def process_slice(self, sl, dim):
        start, stop, step = sl.start, sl.stop, sl.step
        if start is None:
            start = 0
        if start < 0:
            start = self._shape[dim]
        if stop is None:
            stop = self._shape[dim]
        if stop < 0:
            stop = self._shape[dim] + stop
        if step is None:
            step = 1
        assert stop > start, "Start must be less than stop"
        assert step > 0, "No support for  negative increments"
        return slice(start, stop, step)

    def __getitem__(self, idxs):
        """Slices and integers only; an integer keeps its axis with extent 1 (reference lines 460-505)."""
        if not isinstance(idxs, tuple):
            idxs = (idxs,)
        norm = [self.process_slice(s, i) if isinstance(s, slice) else slice(s, s + 1, 1) for i, s in enumerate(idxs)]
        assert len(norm) == self.ndim, "Need indexes equal to number of dimensions"
        shape = [(s.stop - s.start + s.step - 1) // s.step for s in norm]
        offset = self._offset + sum(s.start * st for s, st in zip(norm, self._strides))
        strides = [st * s.step for s, st in zip(norm, self._strides)]
        return BackendTensor.make(shape, strides, self._device, self._handle, offset)

    def __setitem__(self, idxs, other):
        view = self.__getitem__(idxs)
        if isinstance(other, BackendTensor):
            assert prod(view._shape) == prod(other._shape)
            self._device.ewise_setitem(other.compact()._handle, view._handle, view._shape, view._strides, view._offset)
        else:
            self._device.scalar_setitem(prod(view._shape), other, view._handle, view._shape, view._strides, view._offset)

    # ---- elementwise ----------------------------------------------------------------------------------
    def _like(self):
        return BackendTensor.make(self._shape, self._strides, self._device, self._device.Array(self._handle.size))

    def ewise_or_scalar(self, other, ewise_func, scalar_func):
        dev = self._device
        if isinstance(other, BackendTensor):
            if other._shape == self._shape and other._strides == self._strides and self.is_dense() and other.is_dense():
                out = self._like()
                ewise_func(self._handle, other._handle, out._handle)
                return out
            if (other._shape == self._shape and len(self._shape) == 4 and (self.is_channels_last() or other.is_channels_last())
                    and not isinstance(self, PendingTensor) and not isinstance(other, PendingTensor)):
                # one operand lives channels-last (an activation or its gradient), the other compact (e.g. the gradient a
                # `mean` sent back): bring the other one over - one copy, and the result stays in the layout every
                # convolution / BatchNorm kernel wants - instead of two copies to compact and a third one back
                a, b = self.channels_last(), other.channels_last()
                out = a._like()
                ewise_func(a._handle, b._handle, out._handle)
                return out
            if self._shape != other._shape:
                other = other.broadcast_to(self._shape)
            out = BackendTensor.make(self._shape, device=dev)
            ewise_func(self.compact()._handle, other.compact()._handle, out._handle)
            return out
        if self.is_dense():
            out = self._like()
            scalar_func(self._handle, other, out._handle)
            return out
        out = BackendTensor.make(self._shape, device=dev)
        scalar_func(self.compact()._handle, other, out._handle)
        return out

    def _rowvec_operand(self, other):
        """(rows, cols) when `other` is a per-channel / per-column vector for this dense tensor."""
        if not (isinstance(other, BackendTensor) and other.is_compact() and self._device.has("add_rowvec")):
            return None
        if self.ndim == 4 and other._shape == (1, self._shape[1], 1, 1) and self.is_channels_last():
            n, c, h, w = self._shape
            return n * h * w, c
        if self.ndim == 2 and other._shape == (1, self._shape[1]) and self.is_compact():
            return self._shape
        return None

    def __add__(self, other):
        if isinstance(self, PendingTensor) or isinstance(other, PendingTensor):
            fused = PendingTensor.try_add(self, other)
            if fused is not None:
                return fused
        rv = self._rowvec_operand(other)
        if rv is not None:
            out = self._like()
            self._device.add_rowvec(self._handle, other._handle, out._handle, rv[0], rv[1])
            return out
        return self.ewise_or_scalar(other, self._device.ewise_add, self._device.scalar_add)

    __radd__ = __add__

    def __sub__(self, other):
        return self + (-other)

    def __rsub__(self, other):
        return other + (-self)

    def __mul__(self, other):
        return self.ewise_or_scalar(other, self._device.ewise_mul, self._device.scalar_mul)

    __rmul__ = __mul__

    def __truediv__(self, other):
        return self.ewise_or_scalar(other, self._device.ewise_div, self._device.scalar_div)

    def __neg__(self):
        return self * (-1)

    def __pow__(self, other):
        src = self if self.is_dense() else self.compact()
        out = src._like()
        self._device.scalar_power(src._handle, other, out._handle)
        return out

    def maximum(self, other):
        return self.ewise_or_scalar(other, self._device.ewise_maximum, self._device.scalar_maximum)

    def __eq__(self, other):
        return self.ewise_or_scalar(other, self._device.ewise_eq, self._device.scalar_eq)

    def __ge__(self, other):
        return self.ewise_or_scalar(other, self._device.ewise_ge, self._device.scalar_ge)

    def __ne__(self, other):
        return 1 - (self == other)

    def __lt__(self, other):
        return 1 - (self >= other)

    def __gt__(self, other):
        return (self >= other) * (self != other)

    def __le__(self, other):
        return 1 - (self > other)

    __hash__ = object.__hash__

    def _unary(self, fn):
        src = self if self.is_dense() else self.compact()
        out = src._like()
        fn(src._handle, out._handle)
        return out

    def log(self):
        return self._unary(self._device.ewise_log)

    def exp(self):
        return self._unary(self._device.ewise_exp)

    def tanh(self):
        return self._unary(self._device.ewise_tanh)

    # ---- contractions / reductions --------------------------------------------------------------------
    def __matmul__(self, other):
        assert self.ndim == 2 and other.ndim == 2
        assert self._shape[1] == other._shape[0]
        m, n, p = self._shape[0], self._shape[1], other._shape[1]
        out = BackendTensor.make((m, p), device=self._device)
        dev = self._device
        if dev.has("gemm"):
            # read transposed operands in place instead of compacting them (reference lines 612-622)
            a, ta = _gemm_operand(self)
            b, tb = _gemm_operand(other)
            dev.gemm(a._handle, b._handle, out._handle, m, p, n, ta, tb, a._shape[1], b._shape[1], p, 0, None,
                     precision_mode())
        else:
            dev.matmul(self.compact()._handle, other.compact()._handle, out._handle, m, n, p)
        return out

    def reduce_view_out(self, axis, keepdims=False):
        if isinstance(axis, tuple) and not axis:
            raise ValueError("Empty axis in reduce")
        if axis is None:
            view = self.compact().reshape((1,) * (self.ndim - 1) + (prod(self._shape),))
            out = BackendTensor.make((1,) * (self.ndim if keepdims else 1), device=self._device)
            return view, out
        if isinstance(axis, (tuple, list)):
            assert len(axis) == 1, "Only support reduction over a single axis"
            axis = axis[0]
        view = self.permute(tuple(a for a in range(self.ndim) if a != axis) + (axis,))
        if keepdims:
            out_shape = tuple(1 if i == axis else s for i, s in enumerate(self._shape))
        else:
            out_shape = tuple(s for i, s in enumerate(self._shape) if i != axis)
        return view, BackendTensor.make(out_shape, device=self._device)

    def sum(self, axis=None, keepdims=False):
        view, out = self.reduce_view_out(axis, keepdims=keepdims)
        if self._sum_view_div(view, out, 1.0):
            return out
        self._device.reduce_sum(view.compact()._handle, out._handle, view._shape[-1])
        return out

    def _sum_view_div(self, view, out, divisor):
        """Short reductions (<= 32 elements: pooling windows, the two means of a global average pool) straight from the
        strided view, the division of `mean` folded in: one launch instead of compact + reduce_sum + scalar_div, the same
        additions in the same order (dfb_reduce_sum_view_div). False: not available, the caller composes it."""
        dev = self._device
        if not (dev.has("reduce_sum_view_div") and 1 <= view._shape[-1] <= 32 and view.ndim <= 8 and out.size > 0):
            return False
        dev.reduce_sum_view_div(view._handle, out._handle, view._shape, view._strides, view._offset, float(divisor))
        return True

    def max(self, axis=None, keepdims=False):
        view, out = self.reduce_view_out(axis, keepdims=keepdims)
        self._device.reduce_max(view.compact()._handle, out._handle, view._shape[-1])
        return out

    def mean(self, axis=None, keepdims=False):
        # reference quirk Q3 (lines 659-662): divides by the TOTAL element count, also for one axis
        divisor = prod(self._shape)
        if _fix_mean and axis is not None:
            divisor = self._shape[axis[0] if isinstance(axis, (tuple, list)) else axis]
        if divisor != 0 and self._device.has("reduce_sum_view_div"):
            view, out = self.reduce_view_out(axis, keepdims=keepdims)
            if self._sum_view_div(view, out, divisor):
                return out
        return self.sum(axis, keepdims=keepdims) / divisor

    def flip(self, axes):
        assert len(axes) <= len(self._shape)
        strides = list(self._strides)
        offset = self._offset
        for ax in axes:
            offset += (self._shape[ax] - 1) * self._strides[ax]
            strides[ax] = -strides[ax]
        return BackendTensor.make(self._shape, tuple(strides), self._device, self._handle, offset).compact()

    def pad(self, axes):
        assert len(axes) == len(self._shape)
        new_shape = tuple(lo + hi + n for (lo, hi), n in zip(axes, self._shape))
        out = self._device.full(new_shape, 0)
        out[tuple(slice(lo, lo + n) for (lo, _), n in zip(axes, self._shape))] = self
        return out


class BnApply:
    """One training-mode BatchNorm whose statistics exist (mean_var: the producing convolution's epilogue) and whose
    output has not been written yet. `launch` happens at most a few times (a fused sum does not materialise the plain
    output); the running statistics are updated by the first one only."""
    __slots__ = ("x", "mean_var", "gamma", "beta", "save_mean", "save_invstd", "rmean", "rvar", "momentum", "eps", "applied")

    def __init__(self, x, mean_var, gamma, beta, rmean, rvar, momentum, eps):
        dev = x.device
        c = x.shape[1]
        self.x, self.mean_var, self.gamma, self.beta = x, mean_var, gamma, beta
        self.save_mean, self.save_invstd = dev.Array(c), dev.Array(c)
        self.rmean, self.rvar, self.momentum, self.eps = rmean, rvar, float(momentum), float(eps)
        self.applied = False

    def fwd_tuple(self):
        first = not self.applied
        self.applied = True
        h = lambda t: t._handle if t is not None else None  # noqa: E731
        return (self.x._handle, self.mean_var, h(self.gamma), h(self.beta), self.save_mean, self.save_invstd,
                h(self.rmean) if first else None, h(self.rvar) if first else None, self.momentum, self.eps)

    def bwd_tuple(self):
        h = lambda t: t._handle if t is not None else None  # noqa: E731
        return (self.x._handle, self.save_mean, self.save_invstd, h(self.gamma), h(self.beta))


class PendingTensor(BackendTensor):
    """A channels-last activation that is still an expression: relu?( bn_a(x_a) [+ bn_b(x_b)] [+ residual] ), written by
    ONE dfb_bn_fwd_apply launch when somebody needs the values (`_handle`). `a + b` and relu of a pending tensor extend
    the expression instead of launching, so a residual block's `bn2(...) + identity` (+ ReLU) costs one pass."""
    __slots__ = ("_real", "sides", "residual", "relu")

    def __init__(self, *args, **kwargs):
        raise TypeError("use PendingTensor.defer(...)")

    @staticmethod
    def defer(shape, strides, device, sides, residual=None, relu=False):
        t = BackendTensor.__new__(PendingTensor)
        t._shape = tuple(int(v) for v in shape)
        t._strides = tuple(int(v) for v in strides)
        t._offset = 0
        t._device = device
        t._dense = True
        t._aux = None
        t._real = None
        t.sides, t.residual, t.relu = list(sides), residual, bool(relu)
        return t

    @property
    def _handle(self):
        if self._real is None:
            n, c, h, w = self._shape
            y = self._device.Array(n * c * h * w)
            self._device.bn_fwd_apply(self.sides[0].fwd_tuple(), self.sides[1].fwd_tuple() if len(self.sides) > 1 else None,
                                      self.residual._handle if self.residual is not None else None, y, n * h * w, c, self.relu)
            self._real = y
        return self._real

    @_handle.setter
    def _handle(self, value):
        self._real = value

    @property
    def pending(self):
        return self._real is None

    # layout questions are answered from the strides: asking must not launch anything
    def is_dense(self):
        return True

    def is_compact(self):
        return self._strides == _compact_strides(self._shape)

    def is_channels_last(self):
        return True

    def with_relu(self):
        """relu(self) as a new pending tensor, or None when the expression cannot take it."""
        if self._real is not None or self.relu:
            return None
        return PendingTensor.defer(self._shape, self._strides, self._device, self.sides, self.residual, True)

    @staticmethod
    def try_add(a, b):
        """a + b as a (still pending) extension of a pending operand's expression, or None."""
        if not (isinstance(a, BackendTensor) and isinstance(b, BackendTensor)):
            return None
        if a._device != b._device or a._shape != b._shape or a._strides != b._strides:
            return None
        pa = isinstance(a, PendingTensor) and a.pending and not a.relu and a.residual is None
        pb = isinstance(b, PendingTensor) and b.pending and not b.relu and b.residual is None
        if pa and pb and len(a.sides) == 1 and len(b.sides) == 1:
            return PendingTensor.defer(a._shape, a._strides, a._device, [a.sides[0], b.sides[0]])
        if pb and not pa:
            a, b, pa, pb = b, a, pb, pa   # addition commutes bit for bit
        if pa and not pb:
            if isinstance(b, PendingTensor) and b.pending:
                return None               # b is an expression that cannot be merged: let it materialise first
            if not b.is_dense() or b._offset != 0:
                return None
            return PendingTensor.defer(a._shape, a._strides, a._device, a.sides, b)
        return None


def _gemm_operand(t):
    """(tensor, trans) such that tensor is a compact 2-d array and `trans` says whether it holds the
    transpose of the logical operand."""
    if t.is_compact():
        return t, 0
    tt = t.permute((1, 0))
    if tt.is_compact():
        return tt, 1
    return t.compact(), 0


# ---- module-level helpers (reference lines 692-779) ---------------------------------------------------
def Btensor(a, dtype="float32", device=None):
    return BackendTensor(a, device=device)


def empty(shape, dtype="float32", device=None):
    return (device if device is not None else default_device()).empty(shape, dtype)


def full(shape, fill_value, dtype="float32", device=None):
    return (device if device is not None else default_device()).full(shape, fill_value, dtype)


def zeros(shape, dtype="float32", device=None):
    return full(shape, 0.0, dtype, device)


def ones(shape, dtype="float32", device=None):
    return full(shape, 1.0, dtype, device)


def zeros_like(data):
    return data.device.full(data.shape, 0.0, data.dtype)


def ones_like(data):
    return data.device.full(data.shape, 1.0, data.dtype)


def broadcast_to(array, new_shape):
    return array.broadcast_to(new_shape)


def reshape(array, new_shape):
    return array.reshape(new_shape)


def maximum(a, b):
    return a.maximum(b)


def max(a, axis=None, keepdims=False):  # noqa: A001 - reference name
    return a.max(axis, keepdims)


def log(a):
    return a.log()


def exp(a):
    return a.exp()


def tanh(a):
    return a.tanh()


def flip(a, axes):
    return a.flip(axes)


def summation(a, axis=None, keepdims=False):
    return a.sum(axis=axis, keepdims=keepdims)


def mean(a, axis=None, keepdims=False):
    return a.mean(axis=axis, keepdims=keepdims)


def pad(a, axes):
    return a.pad(axes)


def expand_dims(a, axis):
    if abs(axis) > a.ndim:
        raise ValueError("axis {} is out of bounds for Btensor of dimension {}".format(axis, a.ndim))
    return a.compact().reshape(a.shape[:axis] + (1,) + a.shape[axis:])
