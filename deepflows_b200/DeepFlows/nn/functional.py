"""nn.functional: same function names and argument meaning as the reference's
`DeepFlows/nn/functional.py`, with each op that the reference composes from many tensor ops
implemented as ONE graph node over the device's fused (L1) entry points.

    conv2d          reference: __pad2d + __im2col2d (k*k strided setitems) + permute/compact + naive matmul
                    [functional.py:249-344]  ->  implicit-GEMM fprop / dgrad / wgrad kernels
    max_pool2d      reference: pad + im2col + row max, equality-mask backward [347-374, tensor.py:779-791]
                    ->  one forward kernel, one backward kernel with the same tie semantics
    avg_pool2d      reference raises AttributeError [377-404]  ->  implemented
    cross_entropy   reference: 10 tensor ops [104-115]  ->  fused softmax-CE forward / backward
    batch_norm      reference: 16 tensor ops inside BatchNorm2d.forward [modules/batchnorm.py:30-55]
    relu            reference: maximum(x, 0) [15-16]  ->  same forward kernel, one-kernel backward

Activations flow between these ops as logical (N,C,H,W) views over channels-last memory, which is
also the physical order the reference's conv output has before its final transpose [343-344].
"""
from typing import Optional

from .. import tensor
from ..tensor import Tensor, FusedOperator, UnaryOperator, BinaryOperator
from .. import backend_api
from ..backend.backend_tensor import BackendTensor, precision_mode, get_dgrad_mode

LAYOUT_NCHW, LAYOUT_NHWC = 0, 1
WLAYOUT_KCRS, WLAYOUT_KRSC = 0, 1


def _nhwc_view(handle, n, c, h, w, device):
    return BackendTensor.make((n, c, h, w), (h * w * c, 1, w * c, c), device, handle)


# ------------------------------------------------------------------------------------------------
# linear / activations
# ------------------------------------------------------------------------------------------------
def linear(input: Tensor, weight: Tensor, bias: Optional[Tensor] = None):
    """x @ W + b with W stored (in_features, out_features) like the reference [8-12]."""
    out = input @ weight
    return out + bias if bias is not None else out


class _relu(UnaryOperator):
    def forward(self, x: Tensor):
        src = x.data if x.data.is_dense() else x.data.compact()
        self._x = src
        out = src._like()
        src.device.scalar_maximum(src._handle, 0.0, out._handle)
        return out

    def grad_fn(self, x: Tensor, grad):
        src = self._x
        dev = src.device
        if not dev.has("relu_bwd"):
            return (self.data == src) * grad
        if grad.shape != src.shape or grad.strides != src.strides or not grad.is_dense():
            # bring both to the plain compact layout
            src, grad = src.compact(), grad.compact()
        out = src._like()
        dev.relu_bwd(src._handle, grad._handle, out._handle, src._handle.size)
        return out


def relu(input: Tensor) -> Tensor:
    return _relu(input)


class sigmoid(UnaryOperator):
    """1/(1+exp(-x)), evaluated as 0.5*tanh(x/2)+0.5 (no overflow for large |x|). The reference version
    relies on boolean-mask indexing that BackendTensor does not implement [19-27]."""

    def forward(self, input: Tensor):
        return (input.data * 0.5).tanh() * 0.5 + 0.5

    def grad_fn(self, input: Tensor, grad):
        return self.data * (1 - self.data) * grad


class tanh(UnaryOperator):
    def forward(self, input: Tensor):
        return backend_api.tanh(input.data)

    def grad_fn(self, input: Tensor, grad):
        return (1 - self.data ** 2) * grad


def leaky_relu(input: Tensor, negative_slope: float):
    return tensor.maximum(input, input * negative_slope)


def softmax(input: Tensor, dim=None, keepdims=False):
    dim = 1 if dim is None else dim
    shifted = input - tensor.max(input, dim, True)
    e = tensor.exp(shifted)
    return e / tensor.sum(e, dim, True)


def log_softmax(input: Tensor, dim=None, keepdims=False):
    dim = 1 if dim is None else dim
    shifted = input - tensor.max(input, dim, True)
    return shifted - tensor.log(tensor.sum(tensor.exp(shifted), dim, True))


# ------------------------------------------------------------------------------------------------
# losses
# ------------------------------------------------------------------------------------------------
def _reduce_loss(per_elem: Tensor, reduction: str):
    if reduction == "mean":
        return tensor.mean(per_elem)
    if reduction == "sum":
        return tensor.sum(per_elem)
    assert 0, "reduction must be mean or sum."


def l1_loss(input: Tensor, target: Tensor, reduction: str = "mean"):
    diff = input - target
    return _reduce_loss(tensor.maximum(diff, diff * -1), reduction)


def nll_loss(input: Tensor, target: Tensor, reduction: str = "mean"):
    return _reduce_loss(-input * target, reduction)


def mse_loss(input: Tensor, target: Tensor, reduction: str = "mean"):
    return _reduce_loss(tensor.square(input - target), reduction)


def binary_cross_entropy(input: Tensor, target: Tensor, reduction: str = "mean"):
    raise NotImplementedError("binary_cross_entropy is a stub in the reference as well (functional.py:100-101)")


class _softmax_cross_entropy(FusedOperator):
    """loss = scale * sum_i sum_j -(x_ij - max_i - logsumexp_i) * t_ij, shape (1,)."""

    def __init__(self, logits: Tensor, target: Tensor, scale: float):
        self.scale = float(scale)
        super().__init__(logits, target)

    def forward(self, logits, target):
        x, t = logits.data.compact(), target.data.compact()
        self._x, self._t = x, t
        out = BackendTensor.make((1,), device=x.device)
        x.device.softmax_ce_fwd(x._handle, t._handle, out._handle, x.shape[0], x.shape[1], self.scale)
        return out

    def backward_all(self, grad, needs):
        x, t = self._x, self._t
        dx = None
        if needs[0]:
            dx = BackendTensor.make(x.shape, device=x.device)
            x.device.softmax_ce_bwd(x._handle, t._handle, grad.compact()._handle, dx._handle, x.shape[0], x.shape[1],
                                    self.scale)
        dt = None
        if needs[1]:  # d/dt = -scale * log_softmax(x); rare (targets are constants)
            ls = x - x.max(axis=1, keepdims=True).broadcast_to(x.shape)
            ls = ls - ls.exp().sum(axis=1, keepdims=True).log().broadcast_to(x.shape)
            dt = ls * (-self.scale) * grad.broadcast_to(x.shape)
        return dx, dt

    def release(self):
        self._x = self._t = None


def cross_entropy(input: Tensor, target: Tensor, reduction: str = "mean", dim: int = 1):
    """Stable log-softmax cross entropy against dense (one-hot or smoothed) target rows [104-115]."""
    assert reduction in ("mean", "sum"), "reduction must be mean or sum."
    target = target if isinstance(target, Tensor) else Tensor(target, device=input.device)
    if input.ndim == 2 and dim in (1, -1) and input.device.has("softmax_ce_fwd"):
        return _softmax_cross_entropy(input, target, 1.0 / input.shape[0] if reduction == "mean" else 1.0)
    shifted = input - tensor.max(input, dim, True)
    lse = tensor.log(tensor.sum(tensor.exp(shifted), dim, True))
    nll = -(shifted - lse) * target
    if reduction == "mean":
        return tensor.sum(tensor.sum(nll, dim, True)) * (1.0 / input.shape[0])
    return tensor.sum(nll)


# ------------------------------------------------------------------------------------------------
# convolution
# ------------------------------------------------------------------------------------------------
class _conv2d(FusedOperator):
    def __init__(self, x: Tensor, kernel: Tensor, padding: int, stride: int):
        self.padding, self.stride = int(padding), int(stride)
        super().__init__(x, kernel)

    def forward(self, x, kernel):
        xd, wd = x.data, kernel.data
        dev = xd.device
        # Weights live channels-last (K,R,R,C) on the B200 device - the layout the tensor-core kernels read
        # in place, for fprop and (as the transposed operand) for dgrad. A Parameter that arrives compact
        # (fresh from init / load_state_dict / a checkpoint) is re-laid out once, here; its logical shape,
        # `.numpy()` and every other consumer are unaffected (strides carry the layout).
        if wd.ndim == 4 and wd.is_channels_last():
            w_layout = WLAYOUT_KRSC
        elif (wd.ndim == 4 and dev.name == "cuda" and dev.has("WLAYOUT_KRSC")
              and not isinstance(kernel, (UnaryOperator, BinaryOperator, FusedOperator))):  # a stored tensor, not an op result
            kernel.data = wd = wd.channels_last()
            w_layout = WLAYOUT_KRSC
        else:
            wd, w_layout = wd.compact(), WLAYOUT_KCRS
        n, c, h, w = xd.shape
        k, c2, r, r2 = wd.shape
        assert c == c2 and r == r2, "conv2d: kernel %s does not match input %s" % (wd.shape, xd.shape)
        mode = precision_mode()
        if xd.is_channels_last():
            layout = LAYOUT_NHWC
        elif xd.is_compact() and (mode in (0, 3) or c <= 4):
            layout = LAYOUT_NCHW  # the FFMA / first-layer kernels gather straight from NCHW (network input)
        else:
            xd, layout = xd.channels_last(), LAYOUT_NHWC
        p, s = self.padding, self.stride
        oh, ow = (h + 2 * p - r) // s + 1, (w + 2 * p - r) // s + 1
        self._geom = (n, c, h, w, k, r, p, s)
        self._x, self._layout, self._w, self._mode = xd, layout, wd, mode
        self._wl = (w_layout,) if w_layout else ()  # trailing argument only when it is not the default
        y = dev.Array(n * oh * ow * k)
        ws, ws_n = self._workspace(dev)
        dev.conv2d_fprop(xd._handle, layout, wd._handle, y, n, c, h, w, k, r, p, s, mode, ws, ws_n, *self._wl)
        return _nhwc_view(y, n, k, oh, ow, dev)

    def _workspace(self, dev):
        n_floats = dev.conv2d_workspace_floats(*self._geom)
        return (dev.Array(n_floats), n_floats) if n_floats else (None, 0)

    def backward_all(self, grad, needs):
        n, c, h, w, k, r, p, s = self._geom
        dev = self._x.device
        gy = grad.channels_last()
        ws, ws_n = self._workspace(dev)
        dx = dw = None
        # dgrad and wgrad only share their inputs: when both are needed the wgrad goes to the side stream
        # and the two kernels (neither fills 148 SMs on the small layers) run concurrently
        both = needs[0] and needs[1] and dev.has("side_begin")
        if needs[1]:
            # the weight gradient is produced in the weight's own layout, so optimizer and all-reduce walk
            # the two buffers side by side
            dw = _nhwc_view(dev.Array(k * c * r * r), k, c, r, r, dev) if self._wl else BackendTensor.make((k, c, r, r), device=dev)
            if both:
                dev.side_begin()
            try:
                dev.conv2d_wgrad(self._x._handle, self._layout, gy._handle, dw._handle, n, c, h, w, k, r, p, s,
                                 self._mode, ws, ws_n, *self._wl)
            finally:
                if both:
                    dev.side_end()
        if needs[0]:
            buf = dev.Array(n * h * w * c)
            dmode = 0 if get_dgrad_mode() == "reference" else 1
            dev.conv2d_dgrad(gy._handle, self._w._handle, buf, n, c, h, w, k, r, p, s, self._mode, dmode, ws, ws_n,
                             *self._wl)
            dx = _nhwc_view(buf, n, c, h, w, dev)
        if both:
            dev.side_join()
        return dx, dw

    def release(self):
        self._x = self._w = None


def conv2d(x: Tensor, kernel: Tensor, padding: int = 0, stride: int = 1):
    """2-d convolution, x (N,C,H,W), kernel (K,C,R,R); square kernel, symmetric zero padding, scalar
    stride, no dilation/groups - the reference's contract [316-335]."""
    if not isinstance(x, Tensor):
        x = Tensor(x, device=kernel.device)
    return _conv2d(x, kernel, padding, stride)


# ------------------------------------------------------------------------------------------------
# pooling
# ------------------------------------------------------------------------------------------------
class _pool2d(FusedOperator):
    def __init__(self, x: Tensor, kernel_size: int, is_max: bool):
        self.k, self.is_max = int(kernel_size), is_max
        super().__init__(x)

    def forward(self, x):
        xd = x.data.channels_last()
        n, c, h, w = xd.shape
        k = self.k
        oh, ow = (h - k) // k + 1, (w - k) // k + 1
        dev = xd.device
        y = dev.Array(n * oh * ow * c)
        if self.is_max:
            dev.maxpool2d_fwd(xd._handle, y, None, n, h, w, c, k)
        else:
            dev.avgpool2d_fwd(xd._handle, y, n, h, w, c, k)
        out = _nhwc_view(y, n, c, oh, ow, dev)
        self._x, self._y = xd, out
        return out

    def backward_all(self, grad, needs):
        xd = self._x
        n, c, h, w = xd.shape
        dev = xd.device
        gy = grad.channels_last()
        dx = dev.Array(n * h * w * c)
        if self.is_max:  # every tied maximum receives the gradient (reference semantics)
            dev.maxpool2d_bwd(xd._handle, self._y._handle, gy._handle, dx, n, h, w, c, self.k)
        else:
            dev.avgpool2d_bwd(gy._handle, dx, n, h, w, c, self.k)
        return (_nhwc_view(dx, n, c, h, w, dev),)

    def release(self):
        self._x = self._y = None


def _check_pool(name, kernel_size, stride, padding):
    stride = kernel_size if not stride else stride
    if stride != kernel_size or padding != 0:
        raise NotImplementedError(
            "%s: only non-overlapping windows without padding are implemented (kernel_size == stride, padding == 0);"
            " got kernel_size=%s stride=%s padding=%s" % (name, kernel_size, stride, padding))


def max_pool2d(x: Tensor, kernel_size: int, stride: int, padding=0):
    _check_pool("max_pool2d", kernel_size, stride, padding)
    return _pool2d(x, kernel_size, True)


def avg_pool2d(x: Tensor, kernel_size: int, stride: int, padding=0):
    _check_pool("avg_pool2d", kernel_size, stride, padding)
    return _pool2d(x, kernel_size, False)


# ------------------------------------------------------------------------------------------------
# batch normalisation
# ------------------------------------------------------------------------------------------------
class _batch_norm_train(FusedOperator):
    def __init__(self, x, weight, bias, running_mean, running_var, momentum, eps):
        self._rm, self._rv = running_mean, running_var
        self.momentum, self.eps = float(momentum), float(eps)
        self._affine = weight is not None
        super().__init__(*([x, weight, bias] if self._affine else [x]))

    def forward(self, x, weight=None, bias=None):
        xd = x.data.channels_last()
        n, c, h, w = xd.shape
        dev = xd.device
        rows = n * h * w
        g = weight.data.compact() if weight is not None else None
        b = bias.data.compact() if bias is not None else None
        y = dev.Array(rows * c)
        self._mean, self._invstd = dev.Array(c), dev.Array(c)
        rm = self._rm.data.compact() if self._rm is not None else None
        rv = self._rv.data.compact() if self._rv is not None else None
        dev.bn_fwd_train(xd._handle, g._handle if g is not None else None, b._handle if b is not None else None, y,
                         self._mean, self._invstd, rm._handle if rm is not None else None,
                         rv._handle if rv is not None else None, self.momentum, self.eps, rows, c)
        if rm is not None and rm is not self._rm.data:  # running stats were non-compact views: write back
            self._rm.data, self._rv.data = rm, rv
        self._x, self._g = xd, g
        self._rm = self._rv = None
        return _nhwc_view(y, n, c, h, w, dev)

    def backward_all(self, grad, needs):
        xd = self._x
        n, c, h, w = xd.shape
        dev = xd.device
        gy = grad.channels_last()
        dx = dev.Array(n * h * w * c) if needs[0] else None
        dg = BackendTensor.make((1, c, 1, 1), device=dev) if self._affine and needs[1] else None
        db = BackendTensor.make((1, c, 1, 1), device=dev) if self._affine and needs[2] else None
        dev.bn_bwd(xd._handle, gy._handle, self._g._handle if self._g is not None else None, self._mean, self._invstd,
                   dx, dg._handle if dg is not None else None, db._handle if db is not None else None, n * h * w, c)
        out = [_nhwc_view(dx, n, c, h, w, dev) if dx is not None else None]
        if self._affine:
            out += [dg, db]
        return out

    def release(self):
        self._x = self._g = self._mean = self._invstd = None


def batch_norm(x: Tensor, weight, bias, running_mean, running_var, training: bool, momentum: float, eps: float):
    """BatchNorm2d forward. Training: batch mean / biased variance over (N,H,W), running statistics
    updated in place with the biased variance (reference quirk Q7). Evaluation: running statistics."""
    if training:
        return _batch_norm_train(x, weight, bias, running_mean, running_var, momentum, eps)
    if running_mean is None:
        x_hat = x
    elif not tensor.is_grad_enable() and x.device.has("bn_fwd_eval"):
        xd = x.data.channels_last()
        n, c, h, w = xd.shape
        dev = xd.device
        y = dev.Array(n * h * w * c)
        dev.bn_fwd_eval(xd._handle, weight.data.compact()._handle if weight is not None else None,
                        bias.data.compact()._handle if bias is not None else None,
                        running_mean.data.compact()._handle, running_var.data.compact()._handle, y, eps, n * h * w, c)
        return Tensor(_nhwc_view(y, n, c, h, w, dev))
    else:
        x_hat = (x - running_mean) / (running_var + eps) ** 0.5
    if weight is not None:
        return x_hat * weight + bias
    return x_hat


# ------------------------------------------------------------------------------------------------
# 1-d ops (composed from tensor ops; not on the accelerated path)
# ------------------------------------------------------------------------------------------------
def conv1d(input: Tensor, kernel: Tensor, padding: int = 0, stride: int = 1):
    """Not implemented: the reference's own conv1d calls a non-existent `.swapaxes` [166-191], so no
    script can depend on it; 1-d ops are listed as a later widening step."""
    raise NotImplementedError("conv1d is outside the accelerated path (SURVEY 8f rank 4)")


def max_pool1d(x: Tensor, kernel_size: int, stride: int, padding: int = 0):
    raise NotImplementedError("max_pool1d is outside the accelerated path (SURVEY 8f rank 4)")


def avg_pool1d(x: Tensor, kernel_size: int, stride: int, padding: int = 0):
    raise NotImplementedError("avg_pool1d is outside the accelerated path (SURVEY 8f rank 4)")
