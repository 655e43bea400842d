"""A numpy implementation of the device-module protocol (L0 + fused L1 entry points).

TEST INFRASTRUCTURE ONLY (see oracle/numpy_ops.py). It lets the host-side package
(deepflows_b200/DeepFlows: BackendTensor, autograd tape, nn, optim, dist) run on a machine without a
GPU, and it is the CPU baseline's engine. Tests install it with
`DeepFlows.backend_api.register_numpy_device(oracle.numpy_device)`.

L0 follows the reference's CUDA module function by function (cu = DeepFlows/backend/backend_src/
ndarray_backend_cuda.cu:515-716), including the *correct* strided setitem of cu:178-221 - the
reference's own numpy device gets setitem wrong (bt.py:92-100). L1 functions have the
signatures of deepflows_b200/csrc/pybind_shim.cpp and compute with oracle/numpy_ops.py.
"""
import numpy as np

from . import numpy_ops as ops

F32 = np.float32
MODE_FP32, MODE_TF32, MODE_BF16, MODE_SIMT = 0, 1, 2, 3
LAYOUT_NCHW, LAYOUT_NHWC = 0, 1
DGRAD_REFERENCE, DGRAD_EXACT = 0, 1
__device__name__ = "cpu"

counters = {}  # name -> number of calls (used by tests that check fusion on the host side)


def _count(name):
    counters[name] = counters.get(name, 0) + 1


class Array:
    """Flat float32 buffer (reference: CudaArray, cu:48-83)."""
    __slots__ = ("buf", "size")

    def __init__(self, size):
        self.size = int(size)
        self.buf = np.zeros(self.size, dtype=F32)

    def ptr(self):
        return self.buf.ctypes.data

    def __repr__(self):
        return "<numpy_device.Array size=%d>" % self.size


def _flat(h):
    """Array | (Array, offset) | None -> flat numpy view starting at the offset."""
    if h is None:
        return None
    if isinstance(h, tuple):
        return h[0].buf[int(h[1]):]
    return h.buf


# ---- L0 ---------------------------------------------------------------------------------------------
def fill(out, value):
    _count("fill")
    out.buf[...] = F32(value)


def from_numpy(a, out):
    _count("from_numpy")
    a = np.asarray(a)
    if a.size != out.size:
        raise ValueError("Input numpy array size does not match output CudaArray size")
    out.buf[:] = a.reshape(-1).astype(F32)


def to_numpy(a, shape, strides, offset):
    _count("to_numpy")
    return ops.compact(a.buf, shape, strides, offset).reshape(tuple(int(s) for s in shape))


def compact(a, out, shape, strides, offset):
    _count("compact")
    out.buf[:] = ops.compact(a.buf, shape, strides, offset)


def ewise_setitem(a, out, shape, strides, offset):
    _count("ewise_setitem")
    ops.ewise_setitem(a.buf, out.buf, shape, strides, offset)


def scalar_setitem(size, value, out, shape, strides, offset):
    _count("scalar_setitem")
    if size > out.size:
        raise IndexError("ScalarSetitem: size exceeds out array size")
    ops.scalar_setitem(size, value, out.buf, shape, strides, offset)


def _same(*arrays):
    n = arrays[0].size
    if any(a.size != n for a in arrays):
        raise ValueError("Input arrays must have the same size")


def ewise_add(a, b, out):
    _count("ewise_add"); _same(a, b, out); np.add(a.buf, b.buf, out=out.buf)


def ewise_mul(a, b, out):
    _count("ewise_mul"); _same(a, b, out); np.multiply(a.buf, b.buf, out=out.buf)


def ewise_div(a, b, out):
    _count("ewise_div"); _same(a, b, out)
    with np.errstate(divide="ignore", invalid="ignore"):
        np.divide(a.buf, b.buf, out=out.buf)


def ewise_maximum(a, b, out):
    _count("ewise_maximum"); _same(a, b, out); np.maximum(a.buf, b.buf, out=out.buf)


def ewise_eq(a, b, out):
    _count("ewise_eq"); _same(a, b, out); out.buf[:] = (a.buf == b.buf)


def ewise_ge(a, b, out):
    _count("ewise_ge"); _same(a, b, out); out.buf[:] = (a.buf >= b.buf)


def scalar_add(a, v, out):
    _count("scalar_add"); _same(a, out); np.add(a.buf, F32(v), out=out.buf)


def scalar_mul(a, v, out):
    _count("scalar_mul"); _same(a, out); np.multiply(a.buf, F32(v), out=out.buf)


def scalar_div(a, v, out):
    _count("scalar_div"); _same(a, out)
    if F32(v) == 0:
        raise ValueError("ScalarDiv: division by zero")  # cu:305 (std::domain_error -> ValueError)
    np.divide(a.buf, F32(v), out=out.buf)


def scalar_power(a, v, out):
    _count("scalar_power"); _same(a, out)
    with np.errstate(invalid="ignore", divide="ignore"):
        out.buf[:] = np.power(a.buf, F32(v))


def scalar_maximum(a, v, out):
    _count("scalar_maximum"); _same(a, out); np.maximum(a.buf, F32(v), out=out.buf)


def scalar_eq(a, v, out):
    _count("scalar_eq"); _same(a, out); out.buf[:] = (a.buf == F32(v))


def scalar_ge(a, v, out):
    _count("scalar_ge"); _same(a, out); out.buf[:] = (a.buf >= F32(v))


def ewise_log(a, out):
    _count("ewise_log"); _same(a, out); out.buf[:] = ops.ewise_log(a.buf)


def ewise_exp(a, out):
    _count("ewise_exp"); _same(a, out)
    with np.errstate(over="ignore"):
        np.exp(a.buf, out=out.buf)


def ewise_tanh(a, out):
    _count("ewise_tanh"); _same(a, out); np.tanh(a.buf, out=out.buf)


def matmul(a, b, out, m, n, p):
    """out[M,P] = a[M,N] . b[N,P] (MatmulKernel, cu:443-466)."""
    _count("matmul")
    out.buf[:] = (a.buf.reshape(int(m), int(n)) @ b.buf.reshape(int(n), int(p))).reshape(-1)


def reduce_sum(a, out, reduce_size):
    _count("reduce_sum")
    out.buf[:] = a.buf.reshape(-1, int(reduce_size)).sum(axis=1, dtype=F32)


def reduce_max(a, out, reduce_size):
    _count("reduce_max")
    out.buf[:] = a.buf.reshape(-1, int(reduce_size)).max(axis=1)


# ---- runtime shims -----------------------------------------------------------------------------------
def synchronize():
    pass


def set_matmul_mode(mode):
    pass


def launch_count():
    return sum(counters.values())


def tc_launch_count():
    return 0


# ---- L1 ---------------------------------------------------------------------------------------------
def _nchw(flat, n, c, h, w, layout=LAYOUT_NHWC):
    """logical NCHW array over a flat buffer stored NHWC (or NCHW)."""
    if layout == LAYOUT_NCHW:
        return flat[: n * c * h * w].reshape(n, c, h, w)
    return flat[: n * c * h * w].reshape(n, h, w, c).transpose(0, 3, 1, 2)


def _store_nhwc(flat, arr_nchw):
    n, c, h, w = arr_nchw.shape
    flat[: arr_nchw.size] = np.ascontiguousarray(arr_nchw.transpose(0, 2, 3, 1)).reshape(-1)


def copy(src, dst, n):
    _count("copy"); _flat(dst)[:n] = _flat(src)[:n]


def add_n(a, b, out, n):
    _count("add_n"); _flat(out)[:n] = _flat(a)[:n] + _flat(b)[:n]


def mul_n(a, b, out, n):
    _count("mul_n"); _flat(out)[:n] = _flat(a)[:n] * _flat(b)[:n]


def scale_n(a, v, out, n):
    _count("scale_n"); _flat(out)[:n] = _flat(a)[:n] * F32(v)


def fill_n(out, v, n):
    _count("fill_n"); _flat(out)[:n] = F32(v)


def gemm(A, B, C, M, N, K, ta, tb, lda, ldb, ldc, accumulate, bias, mode):
    _count("gemm")
    a, b, c = _flat(A), _flat(B), _flat(C)
    am = a[: (K if ta else M) * lda].reshape(-1, lda)
    am = am[:K, :M].T if ta else am[:M, :K]
    bm = b[: (N if tb else K) * ldb].reshape(-1, ldb)
    bm = bm[:N, :K].T if tb else bm[:K, :N]
    res = (am @ bm).astype(F32)
    if bias is not None:
        res = res + _flat(bias)[:N]
    cm = c[: M * ldc].reshape(M, ldc)
    cm[:, :N] = cm[:, :N] + res if accumulate else res


def conv2d_workspace_floats(N, C, H, W, K, R, pad, stride):
    return 0


WLAYOUT_KCRS, WLAYOUT_KRSC = 0, 1


def _kcrs(w, K, C, R, w_layout):
    """Logical (K,C,R,R) weights from a flat buffer stored (K,C,R,R) or channels-last (K,R,R,C)."""
    flat = _flat(w)[: K * C * R * R]
    return flat.reshape(K, R, R, C).transpose(0, 3, 1, 2) if w_layout == WLAYOUT_KRSC else flat.reshape(K, C, R, R)


def conv2d_fprop(x, x_layout, w, y, N, C, H, W, K, R, pad, stride, mode, ws, ws_floats, w_layout=WLAYOUT_KCRS):
    _count("conv2d_fprop")
    out = ops.conv2d_fprop(_nchw(_flat(x), N, C, H, W, x_layout), _kcrs(w, K, C, R, w_layout), pad, stride)
    _store_nhwc(_flat(y), out)


def conv2d_dgrad(dy, w, dx, N, C, H, W, K, R, pad, stride, mode, dgrad_mode, ws, ws_floats, w_layout=WLAYOUT_KCRS):
    _count("conv2d_dgrad")
    oh, ow = ops.out_size(H, R, pad, stride), ops.out_size(W, R, pad, stride)
    fn = ops.conv2d_dgrad_reference if dgrad_mode == DGRAD_REFERENCE else ops.conv2d_dgrad_exact
    out = fn(_nchw(_flat(dy), N, K, oh, ow), _kcrs(w, K, C, R, w_layout), (N, C, H, W), pad, stride)
    _store_nhwc(_flat(dx), out)


def conv2d_wgrad(x, x_layout, dy, dw, N, C, H, W, K, R, pad, stride, mode, ws, ws_floats, w_layout=WLAYOUT_KCRS):
    _count("conv2d_wgrad")
    oh, ow = ops.out_size(H, R, pad, stride), ops.out_size(W, R, pad, stride)
    out = ops.conv2d_wgrad(_nchw(_flat(x), N, C, H, W, x_layout), _nchw(_flat(dy), N, K, oh, ow), (K, C, R, R), pad, stride)
    if w_layout == WLAYOUT_KRSC:
        out = out.transpose(0, 2, 3, 1)
    _flat(dw)[: out.size] = np.ascontiguousarray(out).reshape(-1)


def add_rowvec(x, v, y, rows, cols):
    _count("add_rowvec")
    _flat(y)[: rows * cols] = (_flat(x)[: rows * cols].reshape(rows, cols) + _flat(v)[:cols]).reshape(-1)


def colsum(x, out, rows, cols):
    _count("colsum")
    _flat(out)[:cols] = _flat(x)[: rows * cols].reshape(rows, cols).sum(axis=0, dtype=np.float64).astype(F32)


def bn_fwd_train(x, gamma, beta, y, save_mean, save_invstd, rmean, rvar, momentum, eps, rows, C):
    _count("bn_fwd_train")
    xa = _flat(x)[: rows * C].reshape(rows, 1, 1, C).transpose(0, 3, 1, 2)  # (rows, C, 1, 1)
    g = _flat(gamma)[:C] if gamma is not None else None
    b = _flat(beta)[:C] if beta is not None else None
    rm = _flat(rmean)[:C] if rmean is not None else None
    rv = _flat(rvar)[:C] if rvar is not None else None
    out, nrm, nrv, mean, var = ops.bn_fwd_train(xa, g, b, rm, rv, momentum, eps)
    _flat(y)[: rows * C] = out.transpose(0, 2, 3, 1).reshape(-1)
    _flat(save_mean)[:C] = mean.reshape(-1)
    _flat(save_invstd)[:C] = (F32(1) / np.sqrt(var + F32(eps))).reshape(-1)
    if rm is not None:
        rm[:] = nrm.reshape(-1)
        rv[:] = nrv.reshape(-1)


def bn_fwd_eval(x, gamma, beta, rmean, rvar, y, eps, rows, C):
    _count("bn_fwd_eval")
    xa = _flat(x)[: rows * C].reshape(rows, 1, 1, C).transpose(0, 3, 1, 2)
    out = ops.bn_fwd_eval(xa, _flat(gamma)[:C] if gamma is not None else None,
                          _flat(beta)[:C] if beta is not None else None, _flat(rmean)[:C], _flat(rvar)[:C], eps)
    _flat(y)[: rows * C] = out.transpose(0, 2, 3, 1).reshape(-1)


def bn_bwd(x, dy, gamma, save_mean, save_invstd, dx, dgamma, dbeta, rows, C):
    _count("bn_bwd")
    xa = _flat(x)[: rows * C].reshape(rows, C).astype(np.float64)
    ga = _flat(dy)[: rows * C].reshape(rows, C).astype(np.float64)
    mean = _flat(save_mean)[:C].astype(np.float64)
    invstd = _flat(save_invstd)[:C].astype(np.float64)
    xh = (xa - mean) * invstd
    db = ga.sum(axis=0)
    dg = (ga * xh).sum(axis=0)
    if dx is not None:
        g = _flat(gamma)[:C].astype(np.float64) if gamma is not None else 1.0
        _flat(dx)[: rows * C] = (g * invstd * (ga - db / rows - xh * dg / rows)).astype(F32).reshape(-1)
    if dgamma is not None:
        _flat(dgamma)[:C] = dg.astype(F32)
    if dbeta is not None:
        _flat(dbeta)[:C] = db.astype(F32)


# ---- the fused halves (include/dfb200.h: dfb_conv2d_fprop_stats ... dfb_bn_bwd_apply), restated from the ops above:
# statistics = the BatchNorm's own (batchnorm.py:33-42), apply = x_hat * gamma + beta (+ the other branch of the block), ---
def colstats_mean_var(x, rows, C, mean_var):
    _count("colstats_mean_var")
    xa = _flat(x)[: rows * C].reshape(rows, C).astype(np.float64)
    mv = _flat(mean_var)
    mv[:C] = xa.mean(axis=0).astype(F32)
    mv[C:2 * C] = xa.var(axis=0).astype(F32)


def conv2d_fprop_stats(x, x_layout, w, w_layout, y, N, C, H, W, K, R, pad, stride, mode, mean_var):
    _count("conv2d_fprop_stats")
    out = ops.conv2d_fprop(_nchw(_flat(x), N, C, H, W, x_layout), _kcrs(w, K, C, R, w_layout), pad, stride)
    _store_nhwc(_flat(y), out)
    colstats_mean_var(y, out.size // K, K, mean_var)


def _bn_side_apply(side, rows, C):
    x, mean_var, gamma, beta, save_mean, save_invstd, rmean, rvar, momentum, eps = side
    xa = _flat(x)[: rows * C].reshape(rows, C)
    mean, var = _flat(mean_var)[:C].copy(), _flat(mean_var)[C:2 * C].copy()
    invstd = (F32(1) / np.sqrt(var + F32(eps))).astype(F32)
    _flat(save_mean)[:C] = mean
    _flat(save_invstd)[:C] = invstd
    if rmean is not None:
        rm, rv = _flat(rmean), _flat(rvar)
        rm[:C] = rm[:C] * F32(1 - momentum) + mean * F32(momentum)
        rv[:C] = rv[:C] * F32(1 - momentum) + var * F32(momentum)
    g = _flat(gamma)[:C] if gamma is not None else F32(1)
    b = _flat(beta)[:C] if beta is not None else F32(0)
    return ((xa - mean) * (invstd * g) + b).astype(F32)


def bn_fwd_apply(a, b, residual, y, rows, C, relu):
    _count("bn_fwd_apply")
    out = _bn_side_apply(a, rows, C)
    if b is not None:
        out = out + _bn_side_apply(b, rows, C)
    if residual is not None:
        out = out + _flat(residual)[: rows * C].reshape(rows, C)
    if relu:
        out = np.maximum(out, F32(0))
    _flat(y)[: rows * C] = out.astype(F32).reshape(-1)


def _bn_side_value(side, rows, C):
    x, mean, invstd, gamma, beta = side
    g = _flat(gamma)[:C] if gamma is not None else F32(1)
    b = _flat(beta)[:C] if beta is not None else F32(0)
    return ((_flat(x)[: rows * C].reshape(rows, C) - _flat(mean)[:C]) * (_flat(invstd)[:C] * g) + b).astype(F32)


def relu_bwd_bn(a, b, residual, dy, dx, rows, C):
    _count("relu_bwd_bn")
    z = _bn_side_value(a, rows, C)
    if b is not None:
        z = z + _bn_side_value(b, rows, C)
    if residual is not None:
        z = z + _flat(residual)[: rows * C].reshape(rows, C)
    g = _flat(dy)[: rows * C].reshape(rows, C)
    _flat(dx)[: rows * C] = np.where(z >= 0, g, F32(0)).astype(F32).reshape(-1)


def bn_bwd_sums(x, dy, save_mean, save_invstd, dbeta, dgamma, rows, C):
    _count("bn_bwd_sums")
    bn_bwd(x, dy, None, save_mean, save_invstd, None, dgamma, dbeta, rows, C)


def bn_bwd_apply(x, dy, gamma, save_mean, save_invstd, dbeta, dgamma, dx, rows, C):
    _count("bn_bwd_apply")
    xa = _flat(x)[: rows * C].reshape(rows, C).astype(np.float64)
    ga = _flat(dy)[: rows * C].reshape(rows, C).astype(np.float64)
    mean, invstd = _flat(save_mean)[:C].astype(np.float64), _flat(save_invstd)[:C].astype(np.float64)
    db, dg = _flat(dbeta)[:C].astype(np.float64), _flat(dgamma)[:C].astype(np.float64)
    g = _flat(gamma)[:C].astype(np.float64) if gamma is not None else 1.0
    _flat(dx)[: rows * C] = (g * invstd * (ga - db / rows - (xa - mean) * invstd * dg / rows)).astype(F32).reshape(-1)


def conv2d_dgrad_fused(dy, w, w_layout, dx, N, C, H, W, K, R, pad, stride, mode, dgrad_mode, addend, bn0, bn1, sums, relu, relu_res):
    _count("conv2d_dgrad_fused")
    conv2d_dgrad(dy, w, dx, N, C, H, W, K, R, pad, stride, mode, dgrad_mode, None, 0, w_layout)
    n = N * C * H * W
    if addend is not None:
        _flat(dx)[:n] = _flat(dx)[:n] + _flat(addend)[:n]
    rows = N * H * W
    if relu:
        relu_bwd_bn(bn0, bn1, relu_res, dx, dx, rows, C)
    for i, bn in enumerate([b for b in (bn0, bn1) if b is not None]):
        bn_bwd_sums(bn[0], dx, bn[1], bn[2], (sums, 0) if not isinstance(sums, tuple) else sums, _off(sums, (1 + i) * C), rows, C)


def _off(h, k):
    return (h[0], h[1] + k) if isinstance(h, tuple) else (h, k)


def dropout_mask(mask, n, keep_prob, state):
    _count("dropout_mask")
    st = _flat(state)
    _flat(mask)[:n] = ops.philox_dropout_mask(n, keep_prob, int(round(float(st[0]))), int(round(float(st[1]))))


def relu_fwd(x, y, n):
    _count("relu_fwd"); _flat(y)[:n] = ops.relu_fwd(_flat(x)[:n])


def relu_bwd(x, dy, dx, n):
    _count("relu_bwd"); _flat(dx)[:n] = ops.relu_bwd(_flat(x)[:n], _flat(dy)[:n])


def maxpool2d_fwd(x, y, idx, N, H, W, C, k):
    _count("maxpool2d_fwd")
    xa = _nchw(_flat(x), N, C, H, W)
    _store_nhwc(_flat(y), ops.maxpool2d_fwd(xa, k))
    if idx is not None:
        am = ops.maxpool2d_argmax(xa, k)
        _flat(idx)[: am.size] = np.ascontiguousarray(am.transpose(0, 2, 3, 1)).reshape(-1).view(F32)


def maxpool2d_bwd(x, y, dy, dx, N, H, W, C, k):
    _count("maxpool2d_bwd")
    oh, ow = (H - k) // k + 1, (W - k) // k + 1
    out = ops.maxpool2d_bwd(_nchw(_flat(x), N, C, H, W), _nchw(_flat(y), N, C, oh, ow), _nchw(_flat(dy), N, C, oh, ow), k)
    _store_nhwc(_flat(dx), out)


def avgpool2d_fwd(x, y, N, H, W, C, k):
    _count("avgpool2d_fwd")
    _store_nhwc(_flat(y), ops.avgpool2d_fwd(_nchw(_flat(x), N, C, H, W), k))


def avgpool2d_bwd(dy, dx, N, H, W, C, k):
    _count("avgpool2d_bwd")
    oh, ow = (H - k) // k + 1, (W - k) // k + 1
    _store_nhwc(_flat(dx), ops.avgpool2d_bwd(_nchw(_flat(dy), N, C, oh, ow), (N, C, H, W), k))


def softmax_ce_fwd(logits, target, loss, rows, cols, scale):
    _count("softmax_ce_fwd")
    _flat(loss)[:1] = ops.softmax_ce_fwd(_flat(logits)[: rows * cols].reshape(rows, cols),
                                         _flat(target)[: rows * cols].reshape(rows, cols), scale)


def softmax_ce_bwd(logits, target, upstream, dlogits, rows, cols, scale):
    _count("softmax_ce_bwd")
    up = _flat(upstream)[0] if upstream is not None else 1.0
    _flat(dlogits)[: rows * cols] = ops.softmax_ce_bwd(_flat(logits)[: rows * cols].reshape(rows, cols),
                                                       _flat(target)[: rows * cols].reshape(rows, cols), up, scale).reshape(-1)


def multi_adam_step(params, grads, m1, m2, sizes, lr, beta1, beta2, eps, weight_decay, t, grad_scale):
    _count("multi_adam_step")
    for p, g, a, b, n in zip(params, grads, m1, m2, sizes):
        pf, gf, af, bf = _flat(p), _flat(g), _flat(a), _flat(b)
        pf[:n], af[:n], bf[:n] = ops.adam_step(pf[:n], gf[:n], af[:n], bf[:n], lr, beta1, beta2, eps, weight_decay, t, grad_scale)


def multi_sgd_step(params, grads, vel, sizes, lr, momentum, weight_decay, nesterov, grad_scale):
    _count("multi_sgd_step")
    for p, g, v, n in zip(params, grads, vel, sizes):
        pf, gf, vf = _flat(p), _flat(g), _flat(v)
        newp, newv = ops.sgd_step(pf[:n], gf[:n], vf[:n] if vf is not None else None, lr, momentum, weight_decay, nesterov, grad_scale)
        pf[:n] = newp
        if momentum > 0.0:
            vf[:n] = newv
