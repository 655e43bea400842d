"""Fused epilogues on a B200 (include/dfb200.h: dfb_conv2d_fprop_stats, dfb_conv2d_dgrad_fused, dfb_bn_fwd_apply,
dfb_relu_bwd_bn, dfb_bn_bwd_sums / dfb_bn_bwd_apply) against the unfused entry points and numpy.

The convolution shapes are the ResNet-18/CIFAR layers of the benchmark (row-halo 32-channel tiles, 64-wide tiles, the
cluster split-K tiles of layers 3-4, stride-2 dgrad with its four parity classes) plus ragged ones (pixel counts that do
not fill a tile, several column tiles)."""
import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu
F32 = np.float32

CONV_SHAPES = [  # N, C, H, W, K, R, pad, stride
    (64, 32, 16, 16, 32, 3, 1, 1),     # layer 1: row-halo kernel, 32-wide tile
    (64, 32, 16, 16, 64, 3, 1, 2),     # layer 2 entry, stride 2
    (64, 64, 8, 8, 64, 3, 1, 1),       # layer 2
    (64, 32, 16, 16, 64, 1, 0, 2),     # 1x1 stride-2 shortcut
    (256, 128, 4, 4, 128, 3, 1, 1),    # layer 3: cluster split-K
    (256, 256, 2, 2, 256, 3, 1, 1),    # layer 4: two column tiles, split 8
    (256, 128, 4, 4, 256, 3, 1, 2),    # layer 4 entry
    (3, 8, 7, 9, 12, 3, 1, 1),         # ragged: 189 pixels, 12 output channels
    (10, 16, 20, 20, 272, 3, 1, 1),    # three column tiles, last one partial
    (256, 32, 16, 16, 32, 3, 1, 1),    # 512 tiles of 32 channels: the persistent kernel (row-halo), two tiles per CTA
    (150, 32, 32, 32, 32, 1, 0, 1),    # 1200 tiles, 1x1 (the first layer's column-matrix convolution), uneven tiles per CTA
    (301, 8, 12, 12, 24, 3, 1, 1),     # persistent, ragged: 339 tiles, the last one partial, 8 -> 24 channels
    (40, 160, 32, 32, 272, 3, 1, 1),   # 320 pixel tiles x 128 x 256: the wide persistent kernel (tc_wide_kernel), partial column tiles
]


def _dev(m, a):
    h = m.Array(a.size)
    m.from_numpy(np.ascontiguousarray(a, dtype=F32).reshape(-1), h)
    return h


def _host(m, h, shape):
    strides = [1] * len(shape)
    for i in range(len(shape) - 2, -1, -1):
        strides[i] = strides[i + 1] * shape[i + 1]
    return m.to_numpy(h, list(shape), strides, 0)


@pytest.mark.parametrize("mode_name", ["tf32", "fp32"])
@pytest.mark.parametrize("geom", CONV_SHAPES)
def test_conv_fprop_stats(cuda_device, geom, mode_name):
    m = cuda_device.mod
    n, c, h, w, k, r, p, s = geom
    mode = m.MODE_TF32 if mode_name == "tf32" else m.MODE_FP32
    oh, ow = (h + 2 * p - r) // s + 1, (w + 2 * p - r) // s + 1
    rng = np.random.RandomState(1)
    x = (rng.randn(n, h, w, c) + 0.5).astype(F32)              # channels-last, non-zero mean
    wt = (rng.randn(k, r, r, c) / np.sqrt(c * r * r)).astype(F32)
    wt[0] += 0.3                                                # one channel with |mean| >> std
    hx, hw = _dev(m, x), _dev(m, wt)
    y0, y1, mv = m.Array(n * oh * ow * k), m.Array(n * oh * ow * k), m.Array(2 * k)
    m.conv2d_fprop(hx, m.LAYOUT_NHWC, hw, y0, n, c, h, w, k, r, p, s, mode, None, 0, m.WLAYOUT_KRSC)
    m.conv2d_fprop_stats(hx, m.LAYOUT_NHWC, hw, m.WLAYOUT_KRSC, y1, n, c, h, w, k, r, p, s, mode, mv)
    a0, a1 = _host(m, y0, (n * oh * ow, k)), _host(m, y1, (n * oh * ow, k))
    assert np.array_equal(a0, a1), "the fused epilogue changed the convolution's output"
    got = _host(m, mv, (2, k))
    a64 = a1.astype(np.float64)
    assert np.abs(got[0] - a64.mean(0)).max() <= 2e-6 * max(1.0, np.abs(a64.mean(0)).max())
    assert rel_err(got[1], a64.var(0)) < 2e-5


@pytest.mark.parametrize("n_bn", [0, 1, 2])
@pytest.mark.parametrize("mode_name", ["tf32", "fp32"])
@pytest.mark.parametrize("geom", CONV_SHAPES)
def test_conv_dgrad_fused(cuda_device, geom, mode_name, n_bn):
    m = cuda_device.mod
    n, c, h, w, k, r, p, s = geom
    if s == 2 and (h % 2 or w % 2):
        pytest.skip("odd stride-2 geometry")
    mode = m.MODE_TF32 if mode_name == "tf32" else m.MODE_FP32
    oh, ow = (h + 2 * p - r) // s + 1, (w + 2 * p - r) // s + 1
    rng = np.random.RandomState(2)
    dy = rng.randn(n, oh, ow, k).astype(F32)
    wt = (rng.randn(k, r, r, c) / np.sqrt(c * r * r)).astype(F32)
    addend = rng.randn(n, h, w, c).astype(F32)
    bx = [(rng.randn(n, h, w, c) * (1 + i) + i).astype(F32) for i in range(2)]
    mean = [b.reshape(-1, c).mean(0).astype(F32) for b in bx]
    invstd = [(1.0 / np.sqrt(b.reshape(-1, c).var(0) + 1e-5)).astype(F32) for b in bx]
    hdy, hw, hadd = _dev(m, dy), _dev(m, wt), _dev(m, addend)
    hbx, hmean, hinv = [_dev(m, b) for b in bx], [_dev(m, v) for v in mean], [_dev(m, v) for v in invstd]
    dx0, dx1, sums = m.Array(n * h * w * c), m.Array(n * h * w * c), m.Array(3 * c)
    m.conv2d_dgrad(hdy, hw, dx0, n, c, h, w, k, r, p, s, mode, m.DGRAD_EXACT, None, 0, m.WLAYOUT_KRSC)
    bns = [(hbx[i], hmean[i], hinv[i], None, None) for i in range(2)]
    m.conv2d_dgrad_fused(hdy, hw, m.WLAYOUT_KRSC, dx1, n, c, h, w, k, r, p, s, mode, m.DGRAD_EXACT, hadd,
                         bns[0] if n_bn > 0 else None, bns[1] if n_bn > 1 else None, sums if n_bn else None, False, None)
    plain = _host(m, dx0, (n * h * w, c))
    got = _host(m, dx1, (n * h * w, c))
    want = plain + addend.reshape(-1, c)
    assert np.array_equal(got, want), "dgrad + addend must equal the separate add bit for bit"
    if n_bn:
        sm = _host(m, sums, (3, c))
        g64 = got.astype(np.float64)
        scale = np.abs(g64).sum(0).max()
        assert np.abs(sm[0] - g64.sum(0)).max() <= 2e-6 * scale
        for i in range(n_bn):
            xh = (bx[i].reshape(-1, c).astype(np.float64) - mean[i]) * invstd[i]
            ref = (g64 * xh).sum(0)
            assert np.abs(sm[1 + i] - ref).max() <= 4e-6 * np.abs(g64 * xh).sum(0).max(), i


@pytest.mark.parametrize("geom", CONV_SHAPES)
def test_lazy_statistics_forward(cuda_device, geom):
    """dfb_conv2d_fprop_stats_lazy -> dfb_bn_fwd_apply: the statistics travel through a statistic slot (fp64 atomics in the
    convolution's epilogue, read in the BatchNorm kernel's prologue, no reduction kernel). Same output, saved statistics
    and running statistics as the eager pair; mean_var holds its floats afterwards; and the slot is clean for its next
    user (the pair is run three times, with a never-consumed lazy call in between)."""
    m = cuda_device.mod
    n, c, h, w, k, r, p, s = geom
    oh, ow = (h + 2 * p - r) // s + 1, (w + 2 * p - r) // s + 1
    rows = n * oh * ow
    rng = np.random.RandomState(11)
    x = (rng.randn(n, h, w, c) + 0.5).astype(F32)
    wt = (rng.randn(k, r, r, c) / np.sqrt(c * r * r)).astype(F32)
    wt[0] += 0.3
    gamma, beta = (rng.rand(k) + 0.5).astype(F32), rng.randn(k).astype(F32)
    hx, hw, hg, hb = _dev(m, x), _dev(m, wt), _dev(m, gamma), _dev(m, beta)

    def run(lazy):
        y, mv, out = m.Array(rows * k), m.Array(2 * k), m.Array(rows * k)
        sm_, si_, rm, rv = m.Array(k), m.Array(k), _dev(m, np.zeros(k, F32)), _dev(m, np.ones(k, F32))
        launches = m.launch_count()
        m.conv2d_fprop_stats(hx, m.LAYOUT_NHWC, hw, m.WLAYOUT_KRSC, y, n, c, h, w, k, r, p, s, m.MODE_TF32, mv, lazy)
        launches = m.launch_count() - launches
        m.bn_fwd_apply((y, mv, hg, hb, sm_, si_, rm, rv, 0.1, 1e-5), None, None, out, rows, k, True)
        return [_host(m, a, sh) for a, sh in ((out, (rows, k)), (mv, (2, k)), (sm_, (k,)), (si_, (k,)), (rm, (k,)), (rv, (k,)))], launches

    want, eager_launches = run(False)
    for rep in range(3):
        got, lazy_launches = run(True)
        assert lazy_launches <= eager_launches, "the lazy call must not launch more kernels than the eager one"
        for name, a, b in zip(("y", "mean_var", "save_mean", "save_invstd", "running_mean", "running_var"), got, want):
            assert np.abs(a.astype(np.float64) - b).max() <= 2e-5 * max(1.0, np.abs(b).max()), (rep, name)
        if rep == 0:   # a lazy producer whose consumer never comes must not poison later users of its slot or key
            y2, mv2 = m.Array(rows * k), m.Array(2 * k)
            m.conv2d_fprop_stats(hx, m.LAYOUT_NHWC, hw, m.WLAYOUT_KRSC, y2, n, c, h, w, k, r, p, s, m.MODE_TF32, mv2, True)
            m.conv2d_fprop_stats(hx, m.LAYOUT_NHWC, hw, m.WLAYOUT_KRSC, y2, n, c, h, w, k, r, p, s, m.MODE_TF32, mv2, False)
            assert np.abs(_host(m, mv2, (2, k)) - want[1]).max() <= 2e-5 * max(1.0, np.abs(want[1]).max())


@pytest.mark.parametrize("n,h,w,c,k", [(8, 16, 16, 32, 2), (3, 9, 11, 12, 2), (2, 12, 12, 64, 3), (5, 7, 7, 5, 2), (64, 32, 32, 32, 2)])
def test_maxpool_relu_bn_backward_fused(cuda_device, n, h, w, c, k):
    """dfb_maxpool_relu_bn_bwd = dfb_maxpool2d_bwd -> dfb_relu_bwd_bn -> dfb_bn_bwd_sums in one pass: the gradient bit for
    bit (ties and the ReLU's z >= 0 rule included: the input is quantised so that windows hold equal maxima and exact
    zeros), the sums to summation-order rounding."""
    m = cuda_device.mod
    rows, oh, ow = n * h * w, (h - k) // k + 1, (w - k) // k + 1
    rng = np.random.RandomState(21)
    x = (np.round(rng.randn(n, h, w, c) * 2) / 2).astype(F32)        # few distinct values: ties
    gamma, beta = np.ones(c, F32), np.zeros(c, F32)
    gamma[::3] = 0.5
    mean, invstd = np.zeros(c, F32), np.ones(c, F32)                 # z = x * gamma exactly: zeros stay zeros
    mean[1::4] = 0.5
    gp = rng.randn(n, oh, ow, c).astype(F32)
    hx, hg, hb, hm, hi, hgp = _dev(m, x), _dev(m, gamma), _dev(m, beta), _dev(m, mean), _dev(m, invstd), _dev(m, gp)
    # forward the way the step runs it: relu(bn(x)) by bn_fwd_apply from given statistics, then the pool
    var = np.zeros(c, F32)   # invstd = 1 / sqrt(var + eps) ~ 1 is not exactly 1: take the kernel's own saved values below
    act, mv, sm_, si_, py_ = m.Array(rows * c), _dev(m, np.concatenate([mean, var])), m.Array(c), m.Array(c), m.Array(n * oh * ow * c)
    m.bn_fwd_apply((hx, mv, hg, hb, sm_, si_, None, None, 0.1, 1e-5), None, None, act, rows, c, True)
    m.maxpool2d_fwd(act, py_, None, n, h, w, c, k)
    # reference chain
    d1, d2, sums_ref = m.Array(rows * c), m.Array(rows * c), m.Array(3 * c)
    m.maxpool2d_bwd(act, py_, hgp, d1, n, h, w, c, k)
    m.relu_bwd_bn((hx, sm_, si_, hg, hb), None, None, d1, d2, rows, c)
    m.bn_bwd_sums(hx, d2, sm_, si_, (sums_ref, 0), (sums_ref, c), rows, c)
    # fused
    d3, sums = m.Array(rows * c), m.Array(3 * c)
    m.maxpool_relu_bn_bwd((hx, sm_, si_, hg, hb), py_, hgp, d3, sums, n, h, w, c, k)
    want, got = _host(m, d2, (rows, c)), _host(m, d3, (rows, c))
    assert np.array_equal(got, want)
    assert np.count_nonzero(want) > 0
    sr, sg = _host(m, sums_ref, (3, c))[:2], _host(m, sums, (3, c))[:2]
    # summation-order rounding, each sum against the sum of the magnitudes of ITS terms: d, and d * x_hat
    xhat = (x.reshape(rows, c) - _host(m, sm_, (c,))) * _host(m, si_, (c,))
    scale0, scale1 = np.abs(want).sum(0).max(), np.abs(want * xhat).sum(0).max()
    assert np.abs(sr[0] - sg[0]).max() <= 4e-6 * max(scale0, 1.0)
    assert np.abs(sr[1] - sg[1]).max() <= 4e-6 * max(scale1, 1.0)


@pytest.mark.parametrize("n_bn", [1, 2])
@pytest.mark.parametrize("geom", [g for g in CONV_SHAPES if not (g[7] == 2 and (g[2] % 2 or g[3] % 2))])
def test_lazy_statistics_backward(cuda_device, geom, n_bn):
    """dfb_conv2d_dgrad_fused_lazy -> n_bn x dfb_bn_bwd_apply: dx of the BatchNorm(s), dbeta and dgamma equal the eager
    chain's; repeated so that the slot is seen to be clean again."""
    m = cuda_device.mod
    n, c, h, w, k, r, p, s = geom
    oh, ow = (h + 2 * p - r) // s + 1, (w + 2 * p - r) // s + 1
    rows = n * h * w
    rng = np.random.RandomState(12)
    dy = rng.randn(n, oh, ow, k).astype(F32)
    wt = (rng.randn(k, r, r, c) / np.sqrt(c * r * r)).astype(F32)
    bx = [(rng.randn(n, h, w, c) * (1 + i) + i).astype(F32) for i in range(2)]
    mean = [b.reshape(-1, c).mean(0).astype(F32) for b in bx]
    invstd = [(1.0 / np.sqrt(b.reshape(-1, c).var(0) + 1e-5)).astype(F32) for b in bx]
    gamma = (rng.rand(c) + 0.5).astype(F32)
    hdy, hw, hg = _dev(m, dy), _dev(m, wt), _dev(m, gamma)
    hbx, hmean, hinv = [_dev(m, b) for b in bx], [_dev(m, v) for v in mean], [_dev(m, v) for v in invstd]
    bns = [(hbx[i], hmean[i], hinv[i], None, None) for i in range(2)]

    def run(lazy):
        dx, sums = m.Array(rows * c), m.Array(3 * c)
        m.conv2d_dgrad_fused(hdy, hw, m.WLAYOUT_KRSC, dx, n, c, h, w, k, r, p, s, m.MODE_TF32, m.DGRAD_EXACT, None,
                             bns[0], bns[1] if n_bn > 1 else None, sums, False, None, lazy)
        outs = []
        for i in range(n_bn):
            dxi = m.Array(rows * c)
            m.bn_bwd_apply(hbx[i], dx, hg, hmean[i], hinv[i], (sums, 0), (sums, (1 + i) * c), dxi, rows, c)
            outs.append(_host(m, dxi, (rows, c)))
        return outs + [_host(m, sums, (3, c))[: 1 + n_bn]]

    want = run(False)
    for rep in range(2):
        got = run(True)
        for i, (a, b) in enumerate(zip(got, want)):
            assert np.abs(a.astype(np.float64) - b).max() <= 2e-5 * max(1.0, np.abs(b).max()), (rep, i)


@pytest.mark.parametrize("dual,res,addend", [(False, False, False), (False, True, True), (True, False, False), (True, True, True)])
@pytest.mark.parametrize("mode_name", ["tf32", "fp32"])
@pytest.mark.parametrize("geom", [CONV_SHAPES[0], CONV_SHAPES[1], CONV_SHAPES[4], CONV_SHAPES[5], CONV_SHAPES[7]])
def test_conv_dgrad_fused_relu(cuda_device, geom, mode_name, dual, res, addend):
    """dgrad whose output is the gradient of relu(bn_0(x_0) [+ bn_1(x_1)] [+ residual]): masked with the recomputed
    pre-activation, then summed - against dgrad (+ addend), dfb_relu_bwd_bn and numpy sums."""
    m = cuda_device.mod
    n, c, h, w, k, r, p, s = geom
    mode = m.MODE_TF32 if mode_name == "tf32" else m.MODE_FP32
    oh, ow = (h + 2 * p - r) // s + 1, (w + 2 * p - r) // s + 1
    rng = np.random.RandomState(6)
    rows = n * h * w
    dy = rng.randn(n, oh, ow, k).astype(F32)
    wt = (rng.randn(k, r, r, c) / np.sqrt(c * r * r)).astype(F32)
    add = rng.randn(rows, c).astype(F32)
    resid = rng.randn(rows, c).astype(F32)
    bx = [(rng.randn(rows, c) * (1 + i) + 0.3 * i).astype(F32) for i in range(2)]
    gam = [(rng.rand(c) + 0.5).astype(F32) for _ in range(2)]
    bet = [(rng.randn(c) * 0.3).astype(F32) for _ in range(2)]
    mean = [b.mean(0).astype(F32) for b in bx]
    invstd = [(1.0 / np.sqrt(b.astype(np.float64).var(0) + 1e-5)).astype(F32) for b in bx]
    hdy, hw, hadd, hres = _dev(m, dy), _dev(m, wt), _dev(m, add), _dev(m, resid)
    bns = [(_dev(m, bx[i]), _dev(m, mean[i]), _dev(m, invstd[i]), _dev(m, gam[i]), _dev(m, bet[i])) for i in range(2)]
    n_bn = 2 if dual else 1
    # separate kernels: dgrad, add, relu backward through the BatchNorm(s)
    d0 = m.Array(rows * c)
    m.conv2d_dgrad(hdy, hw, d0, n, c, h, w, k, r, p, s, mode, m.DGRAD_EXACT, None, 0, m.WLAYOUT_KRSC)
    if addend:
        m.ewise_add(d0, hadd, d0)
    want_h = m.Array(rows * c)
    m.relu_bwd_bn(bns[0], bns[1] if dual else None, hres if res else None, d0, want_h, rows, c)
    want = _host(m, want_h, (rows, c))
    # fused
    dx, sums = m.Array(rows * c), m.Array(3 * c)
    m.conv2d_dgrad_fused(hdy, hw, m.WLAYOUT_KRSC, dx, n, c, h, w, k, r, p, s, mode, m.DGRAD_EXACT, hadd if addend else None,
                         bns[0], bns[1] if dual else None, sums, True, hres if res else None)
    got = _host(m, dx, (rows, c))
    assert np.array_equal(got, want), "fused ReLU mask differs from dfb_relu_bwd_bn"
    sm = _host(m, sums, (3, c))
    g64 = got.astype(np.float64)
    assert np.abs(sm[0] - g64.sum(0)).max() <= 2e-6 * max(np.abs(g64).sum(0).max(), 1e-30)
    for i in range(n_bn):
        xh = (bx[i].astype(np.float64) - mean[i]) * invstd[i]
        assert np.abs(sm[1 + i] - (g64 * xh).sum(0)).max() <= 4e-6 * max(np.abs(g64 * xh).sum(0).max(), 1e-30), i


@pytest.mark.parametrize("relu", [False, True])
@pytest.mark.parametrize("dual", [False, True])
@pytest.mark.parametrize("res", [False, True])
@pytest.mark.parametrize("rows,c", [(4096, 32), (1000, 12), (256, 256), (77, 5)])
def test_bn_fwd_apply_and_relu_bwd(cuda_device, rows, c, res, dual, relu):
    m = cuda_device.mod
    rng = np.random.RandomState(3)
    xs = [(rng.randn(rows, c) * 2 + 1).astype(F32) for _ in range(2)]
    gam = [(rng.rand(c) + 0.5).astype(F32) for _ in range(2)]
    bet = [rng.randn(c).astype(F32) for _ in range(2)]
    r = rng.randn(rows, c).astype(F32)
    eps, mom = 1e-5, 0.1
    y_ref = np.zeros((rows, c))
    sides, keep = [], []
    for i in range(2 if dual else 1):
        hx = _dev(m, xs[i])
        mv = m.Array(2 * c)
        m.colstats_mean_var(hx, rows, c, mv)
        mvh = _host(m, mv, (2, c))
        x64 = xs[i].astype(np.float64)
        assert np.abs(mvh[0] - x64.mean(0)).max() < 1e-5 and rel_err(mvh[1], x64.var(0)) < 2e-5
        sm_, si_, rm, rv = m.Array(c), m.Array(c), _dev(m, np.zeros(c, F32)), _dev(m, np.ones(c, F32))
        sides.append((hx, mv, _dev(m, gam[i]), _dev(m, bet[i]), sm_, si_, rm, rv, mom, eps))
        keep.append((sm_, si_, rm, rv))
        y_ref += (x64 - x64.mean(0)) / np.sqrt(x64.var(0) + eps) * gam[i] + bet[i]
    if res:
        y_ref += r
    if relu:
        y_ref = np.maximum(y_ref, 0)
    hy, hr = m.Array(rows * c), _dev(m, r)
    m.bn_fwd_apply(sides[0], sides[1] if dual else None, hr if res else None, hy, rows, c, relu)
    y = _host(m, hy, (rows, c))
    assert np.abs(y - y_ref).max() <= 2e-5 * max(1.0, np.abs(y_ref).max())
    for i, (sm_, si_, rm, rv) in enumerate(keep):
        x64 = xs[i].astype(np.float64)
        assert np.abs(_host(m, sm_, (c,)) - x64.mean(0)).max() < 1e-5
        assert rel_err(_host(m, si_, (c,)), 1 / np.sqrt(x64.var(0) + eps)) < 2e-5
        assert np.abs(_host(m, rm, (c,)) - mom * x64.mean(0)).max() < 1e-5
        assert rel_err(_host(m, rv, (c,)), (1 - mom) + mom * x64.var(0)) < 2e-5
    if relu:
        dy = rng.randn(rows, c).astype(F32)
        hdy, hdx = _dev(m, dy), m.Array(rows * c)
        five = [(s[0], s[4], s[5], s[2], s[3]) for s in sides]
        m.relu_bwd_bn(five[0], five[1] if dual else None, hr if res else None, hdy, hdx, rows, c)
        dx = _host(m, hdx, (rows, c))
        # the mask is the one the forward kernel applied: wherever it wrote a positive value the gradient passes
        assert np.array_equal(dx[y > 0], dy[y > 0])
        assert np.all((dx == 0) | (dx == dy))
        undecided = np.abs(y_ref) < 1e-4      # pre-activations at rounding distance from 0
        assert np.array_equal(dx[(y == 0) & ~undecided], np.zeros_like(dx)[(y == 0) & ~undecided])


@pytest.mark.parametrize("rows,c", [(4096, 32), (1000, 12), (256, 256)])
def test_bn_bwd_halves_equal_whole(cuda_device, rows, c):
    m = cuda_device.mod
    rng = np.random.RandomState(4)
    x, dy = (rng.randn(rows, c) + 2).astype(F32), rng.randn(rows, c).astype(F32)
    g = (rng.rand(c) + 0.5).astype(F32)
    mean, invstd = x.mean(0).astype(F32), (1 / np.sqrt(x.astype(np.float64).var(0) + 1e-5)).astype(F32)
    hx, hdy, hg, hm, hi = _dev(m, x), _dev(m, dy), _dev(m, g), _dev(m, mean), _dev(m, invstd)
    dx0, dg0, db0 = m.Array(rows * c), m.Array(c), m.Array(c)
    m.bn_bwd(hx, hdy, hg, hm, hi, dx0, dg0, db0, rows, c)
    dx1, sums = m.Array(rows * c), m.Array(2 * c)
    m.bn_bwd_sums(hx, hdy, hm, hi, (sums, 0), (sums, c), rows, c)
    m.bn_bwd_apply(hx, hdy, hg, hm, hi, (sums, 0), (sums, c), dx1, rows, c)
    s = _host(m, sums, (2, c))
    assert rel_err(s[0], _host(m, db0, (c,))) < 1e-5 and rel_err(s[1], _host(m, dg0, (c,))) < 1e-5
    assert rel_err(_host(m, dx1, (rows, c)), _host(m, dx0, (rows, c))) < 1e-5


@pytest.mark.parametrize("mnk", [(1024, 256, 32), (300, 10, 64), (130, 70, 96), (4096, 128, 1152), (128, 4097, 64), (64, 64, 2304)])
@pytest.mark.parametrize("bias,acc", [(False, False), (True, False), (True, True)])
def test_gemm_epilogue(cuda_device, mnk, bias, acc):
    """The coalesced GEMM write-out: row / column tails, bias, accumulate, unaligned leading dimension."""
    m = cuda_device.mod
    M, N, K = mnk
    rng = np.random.RandomState(5)
    A, B = rng.randn(M, K).astype(F32), rng.randn(N, K).astype(F32)
    C0 = rng.randn(M, N).astype(F32)
    b = rng.randn(N).astype(F32)
    hA, hB, hC, hb = _dev(m, A), _dev(m, B), _dev(m, C0), _dev(m, b)
    l0 = m.tc_launch_count()
    m.gemm(hA, hB, hC, M, N, K, 0, 1, K, K, N, 1 if acc else 0, hb if bias else None, m.MODE_TF32)
    if M * N >= 4096:
        assert m.tc_launch_count() > l0
    want = A.astype(np.float64) @ B.astype(np.float64).T + (b if bias else 0) + (C0 if acc else 0)
    got = _host(m, hC, (M, N))
    assert np.abs(got - want).max() <= 2e-2 * np.abs(want).max()
