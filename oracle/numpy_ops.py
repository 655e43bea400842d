"""CPU restatement (numpy) of what the reference's `device='cuda'` training path computes.

TEST INFRASTRUCTURE ONLY. Nothing in the product (deepflows_b200/) imports this package; it is
imported by tests/, by __graft_entry__.smoke() as the checker, and by bench.py's cpu_baseline /
--impl reference legs.

Every function restates the reference algorithm for one op and cites the reference lines it
follows (paths relative to the reference repo; `cu` = DeepFlows/backend/backend_src/
ndarray_backend_cuda.cu, `F.py` = DeepFlows/nn/functional.py, `bt.py` = DeepFlows/backend/
backend_tensor.py). Layout at this level is the reference's logical layout: activations NCHW,
conv weights (K,C,R,R), Linear weights (in,out).

Parity pinning: the reference's own tests hold known answers only for fill / ewise_add / scalar_add
(test/test_cuda.py:47-97; checked in tests/test_oracle.py). Everything else is pinned against outputs
of the reference package itself, imported from /root/reference with its numpy device's two setitem
functions restated from cu:147-221 (the as-shipped numpy setitem is wrong): see
oracle/make_golden.py, which wrote tests/golden/*.npz.
"""
import numpy as np

F32 = np.float32


# ------------------------------------------------------------------------------------------------
# L0: strided gather / scatter over a flat buffer (cu:147-221)
# ------------------------------------------------------------------------------------------------
def view_indices(shape, strides, offset):
    """Flat element index of every view element in row-major view order: cu:147-155 (gid_to_idx)."""
    shape = [int(s) for s in shape]
    idx = np.full(shape if shape else (1,), int(offset), dtype=np.int64)
    for d, (n, st) in enumerate(zip(shape, strides)):
        bshape = [1] * len(shape)
        bshape[d] = n
        idx = idx + (np.arange(n, dtype=np.int64) * int(st)).reshape(bshape)
    return idx.reshape(-1)


def compact(a_flat, shape, strides, offset):
    """out[gid] = a[offset + sum idx_d * stride_d]  (CompactKernel, cu:157-176)."""
    return a_flat[view_indices(shape, strides, offset)].astype(F32)


def ewise_setitem(a_flat, out_flat, shape, strides, offset):
    """out[idx(gid)] = a[gid] for gid < a.size, in gid order so later writes win
    (EwiseSetitemKernel, cu:178-198)."""
    idx = view_indices(shape, strides, offset)[: a_flat.size]
    out_flat[idx] = a_flat[: idx.size]


def scalar_setitem(size, value, out_flat, shape, strides, offset):
    """ScalarSetitemKernel, cu:200-221."""
    out_flat[view_indices(shape, strides, offset)[: int(size)]] = F32(value)


def ewise_log(a):
    """cu:403-414: log of a non-positive number is -inf (numpy would give nan for a < 0)."""
    out = np.full(a.shape, -np.inf, dtype=F32)
    pos = a > 0
    out[pos] = np.log(a[pos])
    return out


# ------------------------------------------------------------------------------------------------
# convolution, F.py:249-344
# ------------------------------------------------------------------------------------------------
def out_size(n, k, pad, stride):
    return (n + 2 * pad - k) // stride + 1


def pad2d(x, p):
    """__pad2d.forward, F.py:302-305 (zero padding)."""
    if p == 0:
        return x
    return np.pad(x, ((0, 0), (0, 0), (p, p), (p, p)))


def im2col2d(xp, k, stride):
    """__im2col2d.forward fallback loop, F.py:275-283: col (N, C, k, k, OH, OW)."""
    n, c, h, w = xp.shape
    oh, ow = (h - k) // stride + 1, (w - k) // stride + 1
    col = np.zeros((n, c, k, k, oh, ow), dtype=F32)
    for i in range(k):
        for j in range(k):
            col[:, :, i, j] = xp[:, :, i:i + oh * stride:stride, j:j + ow * stride:stride]
    return col


def conv2d_fprop(x, w, pad, stride):
    """F.conv2d, F.py:336-344. x (N,C,H,W), w (K,C,R,R) -> (N,K,OH,OW)."""
    n = x.shape[0]
    kout, _, r, _ = w.shape
    col = im2col2d(pad2d(x.astype(F32), pad), r, stride)
    oh, ow = col.shape[-2:]
    a = np.ascontiguousarray(col.transpose(0, 4, 5, 1, 2, 3)).reshape(n * oh * ow, -1)
    out = a @ w.reshape(kout, -1).T.astype(F32)
    return out.reshape(n, oh, ow, kout).transpose(0, 3, 1, 2)


def _dcol(dy, w):
    """d(col matrix) = dY . W, the matmul.grad_fn of tensor.py:699-708 applied to F.py:343, reshaped
    back through F.py:341 to (N, C, k, k, OH, OW)."""
    n, kout, oh, ow = dy.shape
    _, c, r, _ = w.shape
    g = np.ascontiguousarray(dy.transpose(0, 2, 3, 1)).reshape(n * oh * ow, kout).astype(F32)
    dcol = g @ w.reshape(kout, -1).astype(F32)  # (N*OH*OW, C*k*k)
    return dcol.reshape(n, oh, ow, c, r, r).transpose(0, 3, 4, 5, 1, 2)


def conv2d_dgrad_reference(dy, w, x_shape, pad, stride):
    """Input gradient exactly as the reference computes it: __im2col2d.grad_fn ASSIGNS each tap's
    slice (F.py:285-294), so where windows overlap the last (i, j) in loop order wins;
    then __pad2d.grad_fn crops the padding (F.py:307-313)."""
    n, c, h, w_ = x_shape
    r = w.shape[2]
    oh, ow = dy.shape[2], dy.shape[3]
    dcol = _dcol(dy, w)
    gx = np.zeros((n, c, h + 2 * pad, w_ + 2 * pad), dtype=F32)
    for i in range(r):
        for j in range(r):
            gx[:, :, i:i + oh * stride:stride, j:j + ow * stride:stride] = dcol[:, :, i, j]
    return gx[:, :, pad:pad + h, pad:pad + w_] if pad else gx


def conv2d_dgrad_exact(dy, w, x_shape, pad, stride):
    """The true transposed convolution (what `+=` instead of `=` at F.py:292 would give; the 1-d
    version F.py:145 does use `+=`)."""
    n, c, h, w_ = x_shape
    r = w.shape[2]
    oh, ow = dy.shape[2], dy.shape[3]
    dcol = _dcol(dy, w)
    gx = np.zeros((n, c, h + 2 * pad, w_ + 2 * pad), dtype=F32)
    for i in range(r):
        for j in range(r):
            gx[:, :, i:i + oh * stride:stride, j:j + ow * stride:stride] += dcol[:, :, i, j]
    return gx[:, :, pad:pad + h, pad:pad + w_] if pad else gx


def conv2d_wgrad(x, dy, w_shape, pad, stride):
    """Weight gradient: col^T . dY (matmul.grad_fn, tensor.py:709-716) pulled back through the
    reshape / transpose of F.py:342 to (K,C,R,R)."""
    kout, c, r, _ = w_shape
    n, _, oh, ow = dy.shape
    col = im2col2d(pad2d(x.astype(F32), pad), r, stride)
    a = np.ascontiguousarray(col.transpose(0, 4, 5, 1, 2, 3)).reshape(n * oh * ow, -1)
    g = np.ascontiguousarray(dy.transpose(0, 2, 3, 1)).reshape(n * oh * ow, kout).astype(F32)
    return (a.T @ g).T.reshape(kout, c, r, r)


# ------------------------------------------------------------------------------------------------
# pooling, F.py:347-404 + tensor.py:769-791
# ------------------------------------------------------------------------------------------------
def _windows(x, k):
    n, c, h, w = x.shape
    oh, ow = (h - k) // k + 1, (w - k) // k + 1
    xs = x[:, :, :oh * k, :ow * k].reshape(n, c, oh, k, ow, k)
    return xs.transpose(0, 1, 2, 4, 3, 5).reshape(n, c, oh, ow, k * k), oh, ow


def maxpool2d_fwd(x, k):
    """max over each k x k window (F.py:364-374 with stride == k, padding 0)."""
    win, _, _ = _windows(x.astype(F32), k)
    return win.max(axis=-1)


def maxpool2d_argmax(x, k):
    """first arg-max inside the window, row-major (r*k + s), as numpy.argmax."""
    win, _, _ = _windows(x.astype(F32), k)
    return win.argmax(axis=-1).astype(np.int32)


def maxpool2d_bwd(x, y, dy, k):
    """max.grad_fn: (y broadcast == x) * grad, so every tied maximum gets the gradient
    (tensor.py:779-791); rows/cols beyond OH*k get none."""
    n, c, h, w = x.shape
    oh, ow = y.shape[2], y.shape[3]
    dx = np.zeros_like(x, dtype=F32)
    yy = np.repeat(np.repeat(y, k, axis=2), k, axis=3)
    gg = np.repeat(np.repeat(dy, k, axis=2), k, axis=3)
    sub = x[:, :, :oh * k, :ow * k]
    dx[:, :, :oh * k, :ow * k] = np.where(sub == yy, gg, 0).astype(F32)
    return dx


def avgpool2d_fwd(x, k):
    """the arithmetic mean the reference intends at F.py:402 (its code raises)."""
    win, _, _ = _windows(x.astype(F32), k)
    return win.sum(axis=-1, dtype=F32) * F32(1.0 / (k * k))


def avgpool2d_bwd(dy, x_shape, k):
    n, c, h, w = x_shape
    oh, ow = dy.shape[2], dy.shape[3]
    dx = np.zeros(x_shape, dtype=F32)
    dx[:, :, :oh * k, :ow * k] = np.repeat(np.repeat(dy, k, axis=2), k, axis=3) * F32(1.0 / (k * k))
    return dx


# ------------------------------------------------------------------------------------------------
# BatchNorm2d, nn/modules/batchnorm.py:30-55
# ------------------------------------------------------------------------------------------------
def bn_fwd_train(x, gamma, beta, running_mean, running_var, momentum, eps):
    """Returns y, new_running_mean, new_running_var, mean, var. Three chained single-axis sums for
    each statistic (lines 33-42), biased variance, running = run*(1-m) + batch*m (lines 44-46),
    x_hat = (x-mean)/(var+eps)**0.5 (line 47), y = x_hat*gamma + beta (line 53)."""
    x = x.astype(F32)
    n, c, h, w = x.shape
    cnt = F32(n * h * w)
    mean = x.sum(0, keepdims=True).sum(2, keepdims=True).sum(3, keepdims=True) / cnt
    diff = x - mean
    var = (diff * diff).sum(0, keepdims=True).sum(2, keepdims=True).sum(3, keepdims=True) / cnt
    x_hat = diff / (var + F32(eps)) ** F32(0.5)
    y = x_hat
    if gamma is not None:
        y = x_hat * gamma.reshape(1, c, 1, 1) + beta.reshape(1, c, 1, 1)
    m = F32(momentum)
    nrm = nrv = None
    if running_mean is not None:
        nrm = running_mean.reshape(1, c, 1, 1) * (F32(1) - m) + mean * m
        nrv = running_var.reshape(1, c, 1, 1) * (F32(1) - m) + var * m
    return y.astype(F32), nrm, nrv, mean, var


def bn_fwd_eval(x, gamma, beta, running_mean, running_var, eps):
    """lines 49-50, 52-53."""
    c = x.shape[1]
    x_hat = (x.astype(F32) - running_mean.reshape(1, c, 1, 1)) / (running_var.reshape(1, c, 1, 1) + F32(eps)) ** F32(0.5)
    if gamma is not None:
        return x_hat * gamma.reshape(1, c, 1, 1) + beta.reshape(1, c, 1, 1)
    return x_hat


def bn_bwd(x, dy, gamma, eps):
    """Gradient of the composed graph of lines 33-53 w.r.t. x, gamma, beta (what the reference's tape
    produces; equals the textbook batch-norm backward). float64 accumulation for the sums."""
    x64, dy64 = x.astype(np.float64), dy.astype(np.float64)
    n, c, h, w = x.shape
    cnt = n * h * w
    mean = x64.mean(axis=(0, 2, 3), keepdims=True)
    var = x64.var(axis=(0, 2, 3), keepdims=True)
    invstd = 1.0 / np.sqrt(var + eps)
    x_hat = (x64 - mean) * invstd
    dbeta = dy64.sum(axis=(0, 2, 3), keepdims=True)
    dgamma = (dy64 * x_hat).sum(axis=(0, 2, 3), keepdims=True)
    g = gamma.reshape(1, c, 1, 1).astype(np.float64) if gamma is not None else 1.0
    dx = g * invstd * (dy64 - dbeta / cnt - x_hat * dgamma / cnt)
    return dx.astype(F32), dgamma.astype(F32), dbeta.astype(F32)


# ------------------------------------------------------------------------------------------------
# activations / loss
# ------------------------------------------------------------------------------------------------
def relu_fwd(x):
    """F.relu = maximum(x, 0), F.py:15-16."""
    return np.maximum(x, F32(0)).astype(F32)


def relu_bwd(x, dy):
    """maximum.grad_fn: (y == x) * grad, i.e. the gradient passes where x >= 0 (tensor.py:872-877)."""
    return np.where(np.maximum(x, F32(0)) == x, dy, F32(0)).astype(F32)


def softmax_ce_fwd(logits, target, scale):
    """F.cross_entropy, F.py:104-115: m = max; u = x - m; lse = log(sum(exp(u)));
    nll = -(u - lse) * t; mean: sum(sum(nll, dim)) * (1/N). Returns shape (1,)."""
    x = logits.astype(F32)
    u = x - x.max(axis=1, keepdims=True)
    lse = np.log(np.exp(u).sum(axis=1, keepdims=True))
    nll = -(u - lse) * target.astype(F32)
    return np.array([nll.sum(axis=1, keepdims=True).sum() * F32(scale)], dtype=F32)


def softmax_ce_bwd(logits, target, upstream, scale):
    """d loss / d logits = scale * upstream * (softmax * sum_j t - t); the path through the max op
    contributes exactly zero because the softmax row sums to one."""
    x = logits.astype(np.float64)
    e = np.exp(x - x.max(axis=1, keepdims=True))
    sm = e / e.sum(axis=1, keepdims=True)
    t = target.astype(np.float64)
    return (float(scale) * float(upstream) * (sm * t.sum(axis=1, keepdims=True) - t)).astype(F32)


# ------------------------------------------------------------------------------------------------
# optimizers, optim/adam.py:28-63 and optim/sgd.py:16-24 (float32 op by op, scalars rounded to
# float32 where the reference hands them to a scalar_* kernel)
# ------------------------------------------------------------------------------------------------
def adam_step(p, g, v, s, lr, beta1, beta2, eps, weight_decay, t, grad_scale=1.0):
    p, g, v, s = (a.astype(F32) for a in (p, g, v, s))
    if grad_scale != 1.0:
        g = g * F32(grad_scale)
    if weight_decay > 0:
        g = g + p * F32(weight_decay)
    v = v * F32(beta1) + g * F32(1 - beta1)
    s = s * F32(beta2) + (g ** F32(2)) * F32(1 - beta2)
    v_hat = v / F32(1 - beta1 ** t)
    s_hat = s / F32(1 - beta2 ** t)
    upd = v_hat / (s_hat ** F32(0.5) + F32(eps)) * F32(lr)
    return (p + upd * F32(-1)).astype(F32), v.astype(F32), s.astype(F32)


def sgd_step(p, g, vel, lr, momentum, weight_decay, nesterov, grad_scale=1.0):
    p, g = p.astype(F32), g.astype(F32)
    if grad_scale != 1.0:
        g = g * F32(grad_scale)
    g = g + p * F32(weight_decay)
    if momentum > 0.0:
        vel = vel.astype(F32) * F32(momentum) + g
        upd = g + vel * F32(momentum) if nesterov else vel
    else:
        upd = g
    return (p + (upd * F32(lr)) * F32(-1)).astype(F32), vel


# ------------------------------------------------------------------------------------------------
# per-batch preparation of the training scripts (host numpy in the reference):
# augment_batch, test/ResNet_CIFAR10_cuda.py:129-148, and the smoothed one-hot targets, :181-183
# ------------------------------------------------------------------------------------------------
def _reflect_index(i, n):
    """Index into an axis of length n for position i of its numpy 'reflect' padding (edge not repeated)."""
    i = np.abs(i)
    return np.where(i >= n, 2 * (n - 1) - i, i)


def augment_batch(inputs, epoch, num_epochs, pad=4):
    """Same draws from numpy's global generator, in the same order, as the reference function [131-143]; the
    crop / flip of the reflect-padded batch is written as one gather instead of the reference's per-sample slices.
    Pinned against the reference function itself in tests/golden/pipeline.npz (oracle/make_golden_pipeline.py)."""
    n, c, h, w = inputs.shape
    span = 2 * pad + 1
    cy = np.random.randint(0, span, size=n)
    cx = np.random.randint(0, span, size=n)
    flip = np.random.rand(n) < 0.5
    rect = None
    if epoch < num_epochs - 5 and np.random.rand() < 0.2:
        eh = max(1, int(h * np.random.uniform(0.1, 0.2)))
        ew = max(1, int(w * np.random.uniform(0.1, 0.2)))
        rect = (np.random.randint(0, h - eh + 1, size=n), np.random.randint(0, w - ew + 1, size=n), eh, ew)
    hh, ww = np.arange(h), np.arange(w)
    src_h = _reflect_index(cy[:, None] + hh[None, :] - pad, h)                       # (n, h)
    col = np.where(flip[:, None], w - 1 - ww[None, :], ww[None, :])                  # the flip mirrors the crop
    src_w = _reflect_index(cx[:, None] + col - pad, w)                               # (n, w)
    out = inputs[np.arange(n)[:, None, None, None], np.arange(c)[None, :, None, None],
                 src_h[:, None, :, None], src_w[:, None, None, :]]
    if rect is not None:
        ey, ex, eh, ew = rect
        inside = (((hh[None, :] >= ey[:, None]) & (hh[None, :] < ey[:, None] + eh))[:, None, :, None] &
                  ((ww[None, :] >= ex[:, None]) & (ww[None, :] < ex[:, None] + ew))[:, None, None, :])
        out = np.where(inside, inputs.dtype.type(0), out)
    return np.clip(out, -1.0, 1.0)


def smooth_one_hot(labels, num_classes, eps):
    """float32 one-hot rows * (1 - eps) + eps / num_classes, each step rounded to float32 [181-183]."""
    hot = (np.asarray(labels).reshape(-1, 1) == np.arange(num_classes)[None, :]).astype(F32)
    return hot * F32(1 - eps) + F32(eps / num_classes)


def philox_dropout_mask(n, keep_prob, seed, step):
    """Restates dfb_dropout_mask (deepflows_b200/csrc/data_ops.cu): Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel
    random numbers: as easy as 1, 2, 3", SC'11 - the published algorithm; constants 0xD2511F53 / 0xCD9E8D57 / Weyl
    0x9E3779B9, 0xBB67AE85), key (seed, 0xCAFEF00D), counter (i // 4, step, (i // 4) >> 32, 0); element i takes word i % 4;
    mask = float32(word) * 2^-32 < keep_prob. (Opt-in device RNG; the reference itself draws with numpy on the host.)"""
    groups = (n + 3) // 4
    g = np.arange(groups, dtype=np.uint64)
    c = [g & np.uint64(0xFFFFFFFF), np.full(groups, step, np.uint64), g >> np.uint64(32), np.zeros(groups, np.uint64)]
    k0, k1 = np.uint64(seed & 0xFFFFFFFF), np.uint64(0xCAFEF00D)
    m32 = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(0xD2511F53) * c[0]
        p1 = np.uint64(0xCD9E8D57) * c[2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & m32, p1 >> np.uint64(32), p1 & m32
        c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
        k0 = (k0 + np.uint64(0x9E3779B9)) & m32
        k1 = (k1 + np.uint64(0xBB67AE85)) & m32
    words = np.stack(c, axis=1).reshape(-1)[:n].astype(np.uint32)
    u = words.astype(np.float32) * np.float32(2.3283064365386963e-10)
    return (u < np.float32(keep_prob)).astype(np.float32)
