"""L1 parity on a B200: the fused entry points (conv fprop/dgrad/wgrad, GEMM, BatchNorm, ReLU, pooling,
softmax-CE, multi-tensor optimizers) against the oracle on seeded inputs and against the fixtures that
oracle/make_golden.py took from the reference. Integer / index / mask results bit-exact; fp32 mode within
1e-5 relative; TF32 / BF16 operand modes within 2e-2 (north_star tolerances)."""
import numpy as np
import pytest

from conftest import golden, rel_err
from oracle import numpy_ops as ops

pytestmark = pytest.mark.gpu
F32 = np.float32
TOL = {0: 1e-5, 3: 1e-5, 1: 2e-2, 2: 2e-2}  # mode -> relative tolerance


def up(m, a):
    h = m.Array(a.size)
    m.from_numpy(np.ascontiguousarray(a, dtype=F32), h)
    return h


def down(m, h, shape):
    n = int(np.prod(shape))
    return m.to_numpy(h, (n,), (1,), 0).reshape(shape)


def nhwc(a):
    return np.ascontiguousarray(a.transpose(0, 2, 3, 1))


def from_nhwc(a):
    return a.transpose(0, 3, 1, 2)


def run_conv(m, x, w, gy, p, s, mode, x_layout_nchw=False, krsc=False):
    """krsc: weights and weight gradient stored channels-last (K,R,R,C), the layout the host layer keeps conv
    parameters in; the results are returned in logical (K,C,R,R) order either way."""
    n, c, h, wd = x.shape
    k, _, r, _ = w.shape
    oh, ow = ops.out_size(h, r, p, s), ops.out_size(wd, r, p, s)
    ws_n = m.conv2d_workspace_floats(n, c, h, wd, k, r, p, s)
    ws = m.Array(ws_n) if ws_n else None
    hx = up(m, x if x_layout_nchw else nhwc(x))
    layout = m.LAYOUT_NCHW if x_layout_nchw else m.LAYOUT_NHWC
    hw, hgy = up(m, nhwc(w) if krsc else w), up(m, nhwc(gy))
    wl = (m.WLAYOUT_KRSC,) if krsc else ()
    hy, hdx_r, hdx_e, hdw = m.Array(n * oh * ow * k), m.Array(x.size), m.Array(x.size), m.Array(w.size)
    m.conv2d_fprop(hx, layout, hw, hy, n, c, h, wd, k, r, p, s, mode, ws, ws_n, *wl)
    m.conv2d_dgrad(hgy, hw, hdx_r, n, c, h, wd, k, r, p, s, mode, m.DGRAD_REFERENCE, ws, ws_n, *wl)
    m.conv2d_dgrad(hgy, hw, hdx_e, n, c, h, wd, k, r, p, s, mode, m.DGRAD_EXACT, ws, ws_n, *wl)
    m.conv2d_wgrad(hx, layout, hgy, hdw, n, c, h, wd, k, r, p, s, mode, ws, ws_n, *wl)
    dw = from_nhwc(down(m, hdw, (k, r, r, c))) if krsc else down(m, hdw, w.shape)
    return (from_nhwc(down(m, hy, (n, oh, ow, k))), from_nhwc(down(m, hdx_r, (n, h, wd, c))),
            from_nhwc(down(m, hdx_e, (n, h, wd, c))), dw)


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("case", ["k3p1s1", "k5p2s1", "k3p1s2", "k1p0s2", "stem_c3", "k3p0s1_rect", "k3p1s2_odd"])
def test_conv_against_reference_fixture(cuda_device, case, mode):
    m = cuda_device.mod
    g = golden("conv")
    n, c, h, w, k, r, p, s = (int(v) for v in g[case + ".geom"])
    x, wt, gy = g[case + ".x"], g[case + ".w"], g[case + ".gy"]
    y, dx_ref, dx_exact, dw = run_conv(m, x, wt, gy, p, s, mode)
    tol = TOL[mode]
    assert rel_err(y, g[case + ".y"]) < tol
    assert rel_err(dx_ref, g[case + ".dx_ref"]) < tol          # last-writer-wins, F.py:285-294
    assert rel_err(dw, g[case + ".dw"]) < tol
    assert rel_err(dx_exact, ops.conv2d_dgrad_exact(gy, wt, x.shape, p, s)) < tol
    if mode == 0:  # network input arrives NCHW: the FFMA kernels gather from it directly
        y2, _, _, dw2 = run_conv(m, x, wt, gy, p, s, mode, x_layout_nchw=True)
        assert rel_err(y2, g[case + ".y"]) < tol and rel_err(dw2, g[case + ".dw"]) < tol


SHAPES = [  # N, C, H, W, K, R, pad, stride   (scaled-down versions of the C2-C5 layers + ragged edges)
    (5, 3, 32, 32, 64, 3, 1, 1),     # VGG first layer (direct first-layer kernels, K = 64)
    (3, 3, 17, 19, 32, 5, 2, 1),     # CNN-CIFAR10 conv1 (C = 3, 5x5: 75 taps), ragged image
    (2, 2, 9, 9, 8, 3, 0, 2),        # direct kernels: stride 2, no padding
    (2, 4, 8, 8, 12, 3, 1, 1),       # C = 4 but K % 8 != 0: not eligible for the direct kernels
    (8, 32, 16, 16, 32, 3, 1, 1),    # ResNet layer1
    (8, 32, 16, 16, 64, 3, 1, 2),    # layer2 b0.conv1
    (8, 32, 16, 16, 64, 1, 0, 2),    # layer2 shortcut
    (16, 128, 4, 4, 128, 3, 1, 1),   # layer3
    (32, 256, 2, 2, 256, 3, 1, 1),   # layer4
    (4, 1, 28, 28, 32, 5, 2, 1),     # CNN-MNIST conv1 (C = 1)
    (4, 3, 32, 32, 32, 3, 1, 1),     # stem (C = 3)
    (2, 64, 14, 14, 64, 3, 1, 1),    # 224-net tail shape, H not a power of two
    (3, 20, 11, 13, 36, 3, 1, 1),    # ragged everything
    (2, 64, 56, 56, 64, 3, 1, 1),    # VGG-like tile
    (40, 160, 32, 32, 272, 3, 1, 1), # 128 x 256 tiles, more than two per SM: the wide persistent kernel, partial column tiles
]


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("shape", SHAPES)
def test_conv_against_oracle(cuda_device, shape, mode):
    m = cuda_device.mod
    n, c, h, w, k, r, p, s = shape
    rng = np.random.RandomState(sum(shape))
    x = rng.randn(n, c, h, w).astype(F32)
    wt = (rng.randn(k, c, r, r) / np.sqrt(c * r * r)).astype(F32)
    oh, ow = ops.out_size(h, r, p, s), ops.out_size(w, r, p, s)
    gy = rng.randn(n, k, oh, ow).astype(F32)
    y, dx_ref, dx_exact, dw = run_conv(m, x, wt, gy, p, s, mode)
    tol = TOL[mode]
    assert rel_err(y, ops.conv2d_fprop(x, wt, p, s)) < tol
    assert rel_err(dx_ref, ops.conv2d_dgrad_reference(gy, wt, x.shape, p, s)) < tol
    assert rel_err(dx_exact, ops.conv2d_dgrad_exact(gy, wt, x.shape, p, s)) < tol
    assert rel_err(dw, ops.conv2d_wgrad(x, gy, wt.shape, p, s)) < tol


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("shape", SHAPES + [(3, 16, 10, 14, 40, 5, 2, 2), (2, 8, 12, 12, 16, 2, 0, 2), (4, 64, 8, 8, 128, 3, 1, 2),
                                            (2, 160, 8, 8, 24, 3, 1, 2), (64, 256, 2, 2, 256, 3, 1, 1)])
def test_conv_channels_last_weights(cuda_device, shape, mode):
    """Weights / weight gradients stored (K,R,R,C): consumed in place by the tensor-core kernels (fprop as the
    K-major, dgrad as the MN-major operand) and by the FFMA kernels; same results as the (K,C,R,R) layout."""
    m = cuda_device.mod
    n, c, h, w, k, r, p, s = shape
    rng = np.random.RandomState(sum(shape) + 1)
    x = rng.randn(n, c, h, w).astype(F32)
    wt = (rng.randn(k, c, r, r) / np.sqrt(c * r * r)).astype(F32)
    oh, ow = ops.out_size(h, r, p, s), ops.out_size(w, r, p, s)
    gy = rng.randn(n, k, oh, ow).astype(F32)
    t0 = m.tc_launch_count()
    got = run_conv(m, x, wt, gy, p, s, mode, krsc=True)
    tc_launches = m.tc_launch_count() - t0
    want = run_conv(m, x, wt, gy, p, s, mode, krsc=False)
    tol = TOL[mode]
    for name, a, b in zip(("fprop", "dgrad_ref", "dgrad_exact", "wgrad"), got, want):
        assert rel_err(a, b) < (1e-6 if mode == 0 else 1e-3), name  # same arithmetic, possibly another summation order
    assert rel_err(got[0], ops.conv2d_fprop(x, wt, p, s)) < tol
    assert rel_err(got[2], ops.conv2d_dgrad_exact(gy, wt, x.shape, p, s)) < tol
    assert rel_err(got[3], ops.conv2d_wgrad(x, gy, wt.shape, p, s)) < tol
    if mode == 1 and c > 4 and c % 4 == 0 and k % 4 == 0 and (s == 1 or (h % 2 == 0 and w % 2 == 0)):
        assert tc_launches == 3  # fprop, exact dgrad, wgrad all on tcgen05 (the reference-mode dgrad is a gather kernel)


STEM_SHAPES = [  # first layers at training batch sizes (>= 16384 output pixels): image input, C * R * R <= 32
    (16, 3, 32, 32, 32, 3, 1, 1),    # ResNet stem
    (16, 3, 32, 32, 64, 3, 1, 1),    # VGG first layer
    (24, 1, 28, 28, 32, 5, 2, 1),    # CNN-MNIST conv1 (25 taps)
    (72, 2, 31, 33, 16, 3, 0, 2),    # stride 2, odd image, no padding
]


@pytest.mark.parametrize("krsc", [False, True])
@pytest.mark.parametrize("nchw", [False, True])
@pytest.mark.parametrize("shape", STEM_SHAPES)
def test_first_layer_wgrad_at_training_batch(cuda_device, shape, nchw, krsc):
    """TF32 mode, image input from either layout, weights in either layout: the gradient is a column matrix + the
    tcgen05 1x1 wgrad (one tensor-core launch); with DFB_STEM_TC=0 the exact-fp32 gather kernel."""
    import os
    m = cuda_device.mod
    n, c, h, w, k, r, p, s = shape
    rng = np.random.RandomState(sum(shape))
    x = rng.randn(n, c, h, w).astype(F32)
    wt = (rng.randn(k, c, r, r) / np.sqrt(c * r * r)).astype(F32)
    oh, ow = ops.out_size(h, r, p, s), ops.out_size(w, r, p, s)
    gy = rng.randn(n, k, oh, ow).astype(F32)
    t0 = m.tc_launch_count()
    y, _, _, dw = run_conv(m, x, wt, gy, p, s, 1, x_layout_nchw=nchw, krsc=krsc)
    tc_launches = m.tc_launch_count() - t0
    assert rel_err(y, ops.conv2d_fprop(x, wt, p, s)) < TOL[1]
    assert rel_err(dw, ops.conv2d_wgrad(x, gy, wt.shape, p, s)) < TOL[1]
    assert tc_launches == (0 if os.environ.get("DFB_STEM_TC", "1") == "0" else 1)


TC_SHAPES = [  # N, C, H, W, K, R, pad, stride - all TMA-eligible (C, K multiples of 4; stride 2 needs even H, W)
    (8, 32, 16, 16, 32, 3, 1, 1), (8, 32, 16, 16, 64, 3, 1, 2), (8, 32, 16, 16, 64, 1, 0, 2), (16, 128, 4, 4, 128, 3, 1, 1),
    (32, 256, 2, 2, 256, 3, 1, 1), (2, 64, 14, 14, 64, 3, 1, 1), (3, 20, 11, 13, 36, 3, 1, 1), (2, 64, 56, 56, 64, 3, 1, 1),
    (4, 8, 12, 12, 200, 5, 2, 1), (2, 160, 8, 8, 24, 3, 1, 2), (1, 8, 40, 300, 8, 3, 1, 1), (256, 32, 16, 16, 32, 3, 1, 1),
    (3, 16, 10, 14, 40, 5, 2, 2), (2, 8, 12, 12, 16, 2, 0, 2), (4, 64, 8, 8, 128, 3, 1, 2), (2, 12, 6, 6, 8, 4, 1, 2),
    (10, 16, 64, 64, 256, 3, 1, 1), (10, 256, 64, 64, 16, 3, 1, 1),   # enough tiles for the 128 x 256 kernels (fprop / dgrad)
]


@pytest.mark.parametrize("shape", TC_SHAPES)
def test_tf32_conv_runs_on_the_tensor_pipe(cuda_device, shape):
    """TF32 mode: fprop / dgrad / wgrad (stride 1 and 2) must launch tcgen05 kernels, and agree with
    the oracle within the TF32 tolerance (and not better than fp32 rounding: operands really were TF32)."""
    m = cuda_device.mod
    n, c, h, w, k, r, p, s = shape
    rng = np.random.RandomState(sum(shape))
    x = rng.randn(n, c, h, w).astype(F32)
    wt = (rng.randn(k, c, r, r) / np.sqrt(c * r * r)).astype(F32)
    oh, ow = ops.out_size(h, r, p, s), ops.out_size(w, r, p, s)
    gy = rng.randn(n, k, oh, ow).astype(F32)
    hx, hw, hgy = up(m, nhwc(x)), up(m, wt), up(m, nhwc(gy))
    hy, hdx, hdw = m.Array(n * oh * ow * k), m.Array(x.size), m.Array(wt.size)
    t0 = m.tc_launch_count()
    m.conv2d_fprop(hx, m.LAYOUT_NHWC, hw, hy, n, c, h, w, k, r, p, s, m.MODE_TF32, None, 0)
    t1 = m.tc_launch_count()
    m.conv2d_dgrad(hgy, hw, hdx, n, c, h, w, k, r, p, s, m.MODE_TF32, m.DGRAD_EXACT, None, 0)
    t2 = m.tc_launch_count()
    m.conv2d_wgrad(hx, m.LAYOUT_NHWC, hgy, hdw, n, c, h, w, k, r, p, s, m.MODE_TF32, None, 0)
    t3 = m.tc_launch_count()
    assert t1 == t0 + 1 and t3 == t2 + 1 and t2 == t1 + 1
    big = n * oh * ow * k > 50000
    if big:  # float64 im2col of the largest cases is slow; compare against the exact-fp32 kernels instead
        ry, rdx, rdw = m.Array(n * oh * ow * k), m.Array(x.size), m.Array(wt.size)
        m.conv2d_fprop(hx, m.LAYOUT_NHWC, hw, ry, n, c, h, w, k, r, p, s, m.MODE_FP32, None, 0)
        m.conv2d_dgrad(hgy, hw, rdx, n, c, h, w, k, r, p, s, m.MODE_FP32, m.DGRAD_EXACT, None, 0)
        m.conv2d_wgrad(hx, m.LAYOUT_NHWC, hgy, rdw, n, c, h, w, k, r, p, s, m.MODE_FP32, None, 0)
        want = (down(m, ry, (n, oh, ow, k)), down(m, rdx, (n, h, w, c)), down(m, rdw, wt.shape))
    else:
        want = (nhwc(ops.conv2d_fprop(x, wt, p, s)), nhwc(ops.conv2d_dgrad_exact(gy, wt, x.shape, p, s)),
                ops.conv2d_wgrad(x, gy, wt.shape, p, s))
    got = (down(m, hy, (n, oh, ow, k)), down(m, hdx, (n, h, w, c)), down(m, hdw, wt.shape))
    for name, a, b in zip(("fprop", "dgrad", "wgrad"), got, want):
        e = rel_err(a, b)
        assert e < 2e-2, (name, e)
    assert rel_err(got[0], want[0]) > 1e-6


@pytest.mark.parametrize("shape", TC_SHAPES)
def test_fp32_conv_runs_on_the_tensor_pipe(cuda_device, shape):
    """fp32 mode (the package default): fprop / dgrad / wgrad launch the fp32-accurate tcgen05 kernels (operands split into
    TF32 high and low parts, three MMAs per k-step; gemm_tc.cu tc_kernel<P, true>) and agree with the float64 contraction
    (the exact FFMA kernels, MODE_SIMT, for the largest cases) within north_star's 1e-5."""
    m = cuda_device.mod
    n, c, h, w, k, r, p, s = shape
    rng = np.random.RandomState(sum(shape) + 5)
    x = rng.randn(n, c, h, w).astype(F32)
    wt = (rng.randn(k, c, r, r) / np.sqrt(c * r * r)).astype(F32)
    oh, ow = ops.out_size(h, r, p, s), ops.out_size(w, r, p, s)
    gy = rng.randn(n, k, oh, ow).astype(F32)
    hx, hw, hgy = up(m, nhwc(x)), up(m, wt), up(m, nhwc(gy))
    hy, hdx, hdw = m.Array(n * oh * ow * k), m.Array(x.size), m.Array(wt.size)
    t0 = m.tc_launch_count()
    m.conv2d_fprop(hx, m.LAYOUT_NHWC, hw, hy, n, c, h, w, k, r, p, s, m.MODE_FP32, None, 0)
    t1 = m.tc_launch_count()
    m.conv2d_dgrad(hgy, hw, hdx, n, c, h, w, k, r, p, s, m.MODE_FP32, m.DGRAD_EXACT, None, 0)
    t2 = m.tc_launch_count()
    m.conv2d_wgrad(hx, m.LAYOUT_NHWC, hgy, hdw, n, c, h, w, k, r, p, s, m.MODE_FP32, None, 0)
    t3 = m.tc_launch_count()
    assert t1 == t0 + 1 and t3 == t2 + 1 and t2 == t1 + 1
    if n * oh * ow * k > 50000:
        ry, rdx, rdw = m.Array(n * oh * ow * k), m.Array(x.size), m.Array(wt.size)
        m.conv2d_fprop(hx, m.LAYOUT_NHWC, hw, ry, n, c, h, w, k, r, p, s, m.MODE_SIMT, None, 0)
        m.conv2d_dgrad(hgy, hw, rdx, n, c, h, w, k, r, p, s, m.MODE_SIMT, m.DGRAD_EXACT, None, 0)
        m.conv2d_wgrad(hx, m.LAYOUT_NHWC, hgy, rdw, n, c, h, w, k, r, p, s, m.MODE_SIMT, None, 0)
        want = (down(m, ry, (n, oh, ow, k)), down(m, rdx, (n, h, w, c)), down(m, rdw, wt.shape))
    else:
        want = (nhwc(ops.conv2d_fprop(x, wt, p, s)), nhwc(ops.conv2d_dgrad_exact(gy, wt, x.shape, p, s)),
                ops.conv2d_wgrad(x, gy, wt.shape, p, s))
    got = (down(m, hy, (n, oh, ow, k)), down(m, hdx, (n, h, w, c)), down(m, hdw, wt.shape))
    for name, a, b in zip(("fprop", "dgrad", "wgrad"), got, want):
        e = rel_err(a, b)
        assert e < 1e-5, (name, e)


@pytest.mark.parametrize("M,N,K,ta,tb", [(256, 128, 512, 0, 0), (256, 128, 512, 0, 1), (256, 128, 512, 1, 0), (256, 128, 512, 1, 1),
                                         (300, 72, 200, 0, 0), (1000, 260, 36, 1, 0), (8192, 64, 4096, 0, 1), (129, 33, 40, 1, 1),
                                         (19000, 512, 64, 0, 0), (19000, 300, 40, 1, 0)])
def test_fp32_gemm_runs_on_the_tensor_pipe(cuda_device, M, N, K, ta, tb):
    """fp32 mode with TMA-compatible operands: ONE tcgen05 launch (the three-term kernel) and the float64 product within
    1e-5 - and an error three orders of magnitude below the TF32 kernel's on the same operands."""
    m = cuda_device.mod
    rng = np.random.RandomState(M * 7 + N + 1)
    ra, ca = (K, M) if ta else (M, K)
    rb, cb = (N, K) if tb else (K, N)
    lda, ldb = (ca + 3) // 4 * 4, (cb + 3) // 4 * 4
    A, B = rng.randn(ra, lda).astype(F32), rng.randn(rb, ldb).astype(F32)
    a = (A[:, :ca].T if ta else A[:, :ca]).astype(np.float64)
    b = (B[:, :cb].T if tb else B[:, :cb]).astype(np.float64)
    hA, hB, hC, hT = up(m, A), up(m, B), m.Array(M * N), m.Array(M * N)
    before = m.tc_launch_count()
    m.gemm(hA, hB, hC, M, N, K, ta, tb, lda, ldb, N, 0, None, m.MODE_FP32)
    assert m.tc_launch_count() == before + 1
    m.gemm(hA, hB, hT, M, N, K, ta, tb, lda, ldb, N, 0, None, m.MODE_TF32)
    err, err_tf32 = rel_err(down(m, hC, (M, N)), a @ b), rel_err(down(m, hT, (M, N)), a @ b)
    assert err < 1e-5, err
    assert err * 100 < err_tf32, (err, err_tf32)


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("M,N,K,ta,tb", [(64, 10, 3136, 0, 0), (256, 100, 784, 0, 0), (100, 784, 256, 1, 0), (256, 784, 100, 0, 1),
                                         (33, 65, 129, 1, 1), (128, 128, 128, 0, 0), (512, 256, 1024, 0, 1), (4096, 10, 256, 0, 0),
                                         (1000, 36, 77, 1, 0)])
@pytest.mark.parametrize("pad", [(3, 5, 2), (4, 8, 0)])  # odd leading dimensions (FFMA path) and 16-byte aligned ones (TMA path)
def test_gemm_variants(cuda_device, M, N, K, ta, tb, mode, pad):
    m = cuda_device.mod
    rng = np.random.RandomState(M + N + K)
    lda = (M if ta else K) + pad[0]
    ldb = (K if tb else N) + pad[1]
    ldc = N + pad[2]
    A = rng.randn(K if ta else M, lda).astype(F32)
    B = rng.randn(N if tb else K, ldb).astype(F32)
    C0 = rng.randn(M, ldc).astype(F32)
    bias = rng.randn(N).astype(F32)
    a = (A[:K, :M].T if ta else A[:M, :K]).astype(np.float64)
    b = (B[:N, :K].T if tb else B[:K, :N]).astype(np.float64)
    for acc, use_bias in ((0, False), (1, True)):
        hC = up(m, C0)
        m.gemm(up(m, A), up(m, B), hC, M, N, K, ta, tb, lda, ldb, ldc, acc, up(m, bias) if use_bias else None, mode)
        got = down(m, hC, (M, ldc))
        want = a @ b + (bias if use_bias else 0) + (C0[:, :N] if acc else 0)
        assert rel_err(got[:, :N], want) < TOL[mode]
        assert np.array_equal(got[:, N:], C0[:, N:])  # padding columns untouched


@pytest.mark.parametrize("M,N,K,ta,tb", [(256, 128, 512, 0, 0), (256, 128, 512, 0, 1), (256, 128, 512, 1, 0), (256, 128, 512, 1, 1),
                                         (300, 72, 200, 0, 0), (1000, 260, 36, 1, 0), (8192, 64, 4096, 0, 1), (129, 33, 40, 1, 1),
                                         (19000, 512, 64, 0, 0), (19000, 512, 64, 0, 1), (19000, 300, 40, 1, 0), (19000, 512, 64, 1, 1)])
def test_tf32_gemm_runs_on_the_tensor_pipe(cuda_device, M, N, K, ta, tb):
    """TF32 mode with TMA-compatible operands must launch the tcgen05 kernel (no silent FFMA fallback) and
    agree with float64 within the TF32 tolerance; the error must also be ABOVE fp32 rounding, i.e. the
    operands really were consumed as TF32."""
    m = cuda_device.mod
    rng = np.random.RandomState(M * 7 + N)
    ra, ca = (K, M) if ta else (M, K)
    rb, cb = (N, K) if tb else (K, N)
    lda, ldb = (ca + 3) // 4 * 4, (cb + 3) // 4 * 4
    A, B = rng.randn(ra, lda).astype(F32), rng.randn(rb, ldb).astype(F32)
    a = (A[:, :ca].T if ta else A[:, :ca]).astype(np.float64)
    b = (B[:, :cb].T if tb else B[:, :cb]).astype(np.float64)
    hC = m.Array(M * N)
    before = m.tc_launch_count()
    m.gemm(up(m, A), up(m, B), hC, M, N, K, ta, tb, lda, ldb, N, 0, None, m.MODE_TF32)
    assert m.tc_launch_count() == before + 1
    err = rel_err(down(m, hC, (M, N)), a @ b)
    assert err < 2e-2
    assert err > 1e-6


def test_batchnorm_against_reference_fixture(cuda_device):
    m = cuda_device.mod
    g = golden("ops")
    x, gy = g["bn.x"], g["bn.gy"]
    n, c, h, w = x.shape
    rows = n * h * w
    hx, hgy = up(m, nhwc(x)), up(m, nhwc(gy))
    hg, hb = up(m, g["bn.gamma"]), up(m, g["bn.beta"])
    hrm, hrv = up(m, np.zeros(c, F32)), up(m, np.ones(c, F32))
    hy, hmean, hinv = m.Array(x.size), m.Array(c), m.Array(c)
    m.bn_fwd_train(hx, hg, hb, hy, hmean, hinv, hrm, hrv, 0.1, 1e-5, rows, c)
    assert rel_err(from_nhwc(down(m, hy, (n, h, w, c))), g["bn.y"]) < 1e-5
    assert rel_err(down(m, hrm, (1, c, 1, 1)), g["bn.running_mean"]) < 1e-5
    assert rel_err(down(m, hrv, (1, c, 1, 1)), g["bn.running_var"]) < 1e-5
    hdx, hdg, hdb = m.Array(x.size), m.Array(c), m.Array(c)
    m.bn_bwd(hx, hgy, hg, hmean, hinv, hdx, hdg, hdb, rows, c)
    assert rel_err(from_nhwc(down(m, hdx, (n, h, w, c))), g["bn.dx"]) < 5e-5
    assert rel_err(down(m, hdg, (1, c, 1, 1)), g["bn.dgamma"]) < 5e-5
    assert rel_err(down(m, hdb, (1, c, 1, 1)), g["bn.dbeta"]) < 5e-5
    m.bn_fwd_eval(hx, hg, hb, hrm, hrv, hy, 1e-5, rows, c)
    assert rel_err(from_nhwc(down(m, hy, (n, h, w, c))), g["bn.y_eval"]) < 1e-5


@pytest.mark.parametrize("n,c,h,w", [(256, 32, 16, 16), (64, 256, 2, 2), (3, 5, 7, 9), (16, 1024, 3, 3), (2, 3, 2, 2), (1, 4, 1, 1),
                                     (64, 64, 16, 16), (16, 128, 16, 16), (5, 48, 30, 30), (256, 256, 2, 2), (9, 32, 11, 11)])
def test_batchnorm_against_oracle(cuda_device, n, c, h, w):
    m = cuda_device.mod
    rng = np.random.RandomState(n + c)
    x = (rng.randn(n, c, h, w) * rng.rand(1, c, 1, 1) * 3 + rng.randn(1, c, 1, 1) * 10).astype(F32)  # mean >> std
    gy = rng.randn(n, c, h, w).astype(F32)
    gamma, beta = (rng.rand(c) + 0.5).astype(F32), rng.randn(c).astype(F32)
    rm0, rv0 = rng.randn(c).astype(F32), (rng.rand(c) + 0.5).astype(F32)
    rows = n * h * w
    hx, hgy, hg, hb, hrm, hrv = up(m, nhwc(x)), up(m, nhwc(gy)), up(m, gamma), up(m, beta), up(m, rm0), up(m, rv0)
    hy, hmean, hinv, hdx, hdg, hdb = m.Array(x.size), m.Array(c), m.Array(c), m.Array(x.size), m.Array(c), m.Array(c)
    m.bn_fwd_train(hx, hg, hb, hy, hmean, hinv, hrm, hrv, 0.1, 1e-5, rows, c)
    m.bn_bwd(hx, hgy, hg, hmean, hinv, hdx, hdg, hdb, rows, c)
    x64 = x.astype(np.float64)
    mean, var = x64.mean(axis=(0, 2, 3)), x64.var(axis=(0, 2, 3))
    y = (x64 - mean.reshape(1, c, 1, 1)) / np.sqrt(var.reshape(1, c, 1, 1) + 1e-5) * gamma.reshape(1, c, 1, 1) + beta.reshape(1, c, 1, 1)
    # float32 inputs resolve x - mean only to ~0.5 ulp of |mean|: a channel with |mean| >> std (the data above has
    # some with mean/std > 1e4) carries that into y as ulp(mean) * invstd * gamma whatever the implementation does
    cond = (np.abs(mean) * gamma / np.sqrt(var + 1e-5)).max()
    assert np.abs(from_nhwc(down(m, hy, (n, h, w, c))) - y).max() < 2e-5 * max(1.0, np.abs(y).max()) + 4 * 6e-8 * cond
    assert rel_err(down(m, hmean, (c,)), mean) < 1e-6
    assert rel_err(down(m, hrv, (c,)), rv0 * 0.9 + var * 0.1) < 1e-5
    assert rel_err(down(m, hrm, (c,)), rm0 * 0.9 + mean * 0.1) < 1e-5
    dx, dg, db = ops.bn_bwd(x, gy, gamma, 1e-5)
    assert rel_err(from_nhwc(down(m, hdx, (n, h, w, c))), dx) < 1e-4
    assert rel_err(down(m, hdg, (1, c, 1, 1)), dg) < 1e-4 and rel_err(down(m, hdb, (1, c, 1, 1)), db) < 1e-4


def test_relu_and_pool_bit_exact(cuda_device):
    m = cuda_device.mod
    g = golden("ops")
    x, gy = g["relu.x"], g["relu.gy"]
    hx, hy, hdx = up(m, x), m.Array(x.size), m.Array(x.size)
    m.relu_fwd(hx, hy, x.size)
    m.relu_bwd(hx, up(m, gy), hdx, x.size)
    assert np.array_equal(down(m, hy, x.shape), g["relu.y"]) and np.array_equal(down(m, hdx, x.shape), g["relu.dx"])
    for name in ("pool", "pool_odd"):
        x, gy, y = g[name + ".x"], g[name + ".gy"], g[name + ".y"]
        n, c, h, w = x.shape
        oh, ow = y.shape[2:]
        hx, hy, hidx, hdx, hdx2 = up(m, nhwc(x)), m.Array(y.size), m.Array(y.size), m.Array(x.size), m.Array(x.size)
        m.maxpool2d_fwd(hx, hy, hidx, n, h, w, c, 2)
        assert np.array_equal(from_nhwc(down(m, hy, (n, oh, ow, c))), y)
        idx = m.to_numpy_i32(hidx, y.size).reshape(n, oh, ow, c).transpose(0, 3, 1, 2)
        assert np.array_equal(idx, g[name + ".argmax"])                       # argmax indices bit-exact
        m.maxpool2d_bwd(hx, hy, up(m, nhwc(gy)), hdx, n, h, w, c, 2)
        assert np.array_equal(from_nhwc(down(m, hdx, (n, h, w, c))), g[name + ".dx"])  # ties: all maxima get grad
        m.maxpool2d_bwd_idx(hidx, up(m, nhwc(gy)), hdx2, n, h, w, c, 2)
        routed = from_nhwc(down(m, hdx2, (n, h, w, c)))
        assert np.count_nonzero(routed) <= y.size and rel_err(routed.sum(), gy.sum()) < 1e-5
        hya, hdxa = m.Array(y.size), m.Array(x.size)
        m.avgpool2d_fwd(hx, hya, n, h, w, c, 2)
        assert rel_err(from_nhwc(down(m, hya, (n, oh, ow, c))), ops.avgpool2d_fwd(x, 2)) < 1e-6
        m.avgpool2d_bwd(up(m, nhwc(gy)), hdxa, n, h, w, c, 2)
        assert rel_err(from_nhwc(down(m, hdxa, (n, h, w, c))), ops.avgpool2d_bwd(gy, x.shape, 2)) < 1e-6


def test_softmax_ce(cuda_device):
    m = cuda_device.mod
    g = golden("ops")
    for name, scale in (("ce_mean", 1 / 16), ("ce_sum", 1.0)):
        lg, tg = g[name + ".logits"], g[name + ".target"]
        hl, ht, hloss, hd = up(m, lg), up(m, tg), m.Array(1), m.Array(lg.size)
        m.softmax_ce_fwd(hl, ht, hloss, 16, 10, scale)
        m.softmax_ce_bwd(hl, ht, up(m, np.ones(1, F32)), hd, 16, 10, scale)
        assert rel_err(down(m, hloss, (1,)), g[name + ".loss"]) < 1e-5
        assert rel_err(down(m, hd, lg.shape), g[name + ".dlogits"]) < 5e-5
    rng = np.random.RandomState(0)
    lg = (rng.randn(4096, 10) * 5).astype(F32)
    tg = np.eye(10, dtype=F32)[rng.randint(0, 10, 4096)]
    hloss = m.Array(1)
    m.softmax_ce_fwd(up(m, lg), up(m, tg), hloss, 4096, 10, 1 / 4096)
    assert rel_err(down(m, hloss, (1,)), ops.softmax_ce_fwd(lg, tg, 1 / 4096)) < 1e-5


@pytest.mark.parametrize("M,K,N,bias", [(256, 256, 10, True), (64, 20, 10, True), (7, 33, 1, False), (300, 2048, 16, True), (1, 5, 3, False)])
def test_linear_small(cuda_device, M, K, N, bias):
    """dfb_linear_small_fwd / _bwd (the classifier in one launch each way) against the float64 contraction at the fp32-mode
    tolerance of north_star (1e-5 relative), including the partial-output forms of the backward."""
    m = cuda_device.mod
    rng = np.random.RandomState(M + K + N)
    x, w, b, gy = rng.randn(M, K).astype(F32), (rng.randn(K, N) / np.sqrt(K)).astype(F32), rng.randn(N).astype(F32), rng.randn(M, N).astype(F32)
    hx, hw, hb, hg = up(m, x), up(m, w), up(m, b) if bias else None, up(m, gy)
    hy = m.Array(M * N)
    m.linear_small_fwd(hx, hw, hb, hy, M, K, N)
    want = x.astype(np.float64) @ w.astype(np.float64) + (b if bias else 0.0)
    assert rel_err(down(m, hy, (M, N)), want) < 1e-5
    hdx, hdw, hdb = m.Array(M * K), m.Array(K * N), m.Array(N)
    m.linear_small_bwd(hx, hw, hg, hdx, hdw, hdb, M, K, N)
    g64 = gy.astype(np.float64)
    assert rel_err(down(m, hdx, (M, K)), g64 @ w.astype(np.float64).T) < 1e-5
    assert rel_err(down(m, hdw, (K, N)), x.astype(np.float64).T @ g64) < 1e-5
    assert rel_err(down(m, hdb, (N,)), g64.sum(0)) < 1e-5
    # only the weight gradient (frozen input, no bias), only the input gradient
    hdw2, hdx2 = m.Array(K * N), m.Array(M * K)
    m.linear_small_bwd(hx, None, hg, None, hdw2, None, M, K, N)
    m.linear_small_bwd(None, hw, hg, hdx2, None, None, M, K, N)
    assert np.array_equal(down(m, hdw2, (K, N)), down(m, hdw, (K, N))) and np.array_equal(down(m, hdx2, (M, K)), down(m, hdx, (M, K)))
    with pytest.raises(ValueError):
        m.linear_small_fwd(hx, hw, hb, hy, M, K, 17)


def test_rowvec_and_colsum(cuda_device):
    m = cuda_device.mod
    rng = np.random.RandomState(0)
    for rows, cols in ((65536, 32), (1000, 10), (7, 3), (4096, 256), (5, 1028)):
        x, v = rng.randn(rows, cols).astype(F32), rng.randn(cols).astype(F32)
        hy, hs = m.Array(x.size), m.Array(cols)
        m.add_rowvec(up(m, x), up(m, v), hy, rows, cols)
        assert np.array_equal(down(m, hy, x.shape), x + v)
        m.colsum(up(m, x), hs, rows, cols)
        assert np.abs(down(m, hs, (cols,)) - x.astype(np.float64).sum(axis=0)).max() < 1e-5 * np.abs(x).sum(axis=0).max()


def test_multi_tensor_optimizers(cuda_device):
    m = cuda_device.mod
    g = golden("optim")
    for name, kw in (("adam", dict(lr=5e-3, wd=5e-4)), ("adam_nowd", dict(lr=1e-3, wd=0.0))):
        ps = [up(m, g["%s.p0.%d" % (name, i)]) for i in range(3)]
        sizes = [g["%s.p0.%d" % (name, i)].size for i in range(3)]
        vs, ss = [m.Array(n) for n in sizes], [m.Array(n) for n in sizes]
        for a in vs + ss:
            m.fill(a, 0.0)
        for st in range(3):
            gs = [up(m, g["%s.g%d.%d" % (name, st, i)]) for i in range(3)]
            m.multi_adam_step(ps, gs, vs, ss, sizes, kw["lr"], 0.9, 0.999, 1e-8, kw["wd"], st + 1, 1.0)
        for i in range(3):
            want = g["%s.p3.%d" % (name, i)]
            assert rel_err(down(m, ps[i], want.shape), want) < 1e-6
    for name, kw in (("sgd", dict(mom=0.0, wd=0.0, nes=False)), ("sgd_mom", dict(mom=0.9, wd=1e-3, nes=True))):
        ps = [up(m, g["%s.p0.%d" % (name, i)]) for i in range(3)]
        sizes = [g["%s.p0.%d" % (name, i)].size for i in range(3)]
        vs = [m.Array(n) for n in sizes]
        for a in vs:
            m.fill(a, 0.0)
        for st in range(3):
            gs = [up(m, g["%s.g%d.%d" % (name, st, i)]) for i in range(3)]
            m.multi_sgd_step(ps, gs, vs, sizes, 0.05, kw["mom"], kw["wd"], kw["nes"], 1.0)
        for i in range(3):
            want = g["%s.p3.%d" % (name, i)]
            assert rel_err(down(m, ps[i], want.shape), want) < 1e-6
    # many tensors of awkward sizes, one launch; compare with the oracle's op-by-op float32 Adam
    rng = np.random.RandomState(1)
    sizes = [1, 3, 4, 5, 4095, 4096, 4097, 10000, 123457]
    p0 = [rng.randn(n).astype(F32) for n in sizes]
    g0 = [rng.randn(n).astype(F32) for n in sizes]
    ps, gs = [up(m, p) for p in p0], [up(m, x) for x in g0]
    vs, ss = [m.Array(n) for n in sizes], [m.Array(n) for n in sizes]
    for a in vs + ss:
        m.fill(a, 0.0)
    before = m.launch_count()
    m.multi_adam_step(ps, gs, vs, ss, sizes, 1e-3, 0.9, 0.999, 1e-8, 5e-4, 1, 0.5)
    assert m.launch_count() - before == 1
    for i, n in enumerate(sizes):
        want, _, _ = ops.adam_step(p0[i], g0[i], np.zeros(n, F32), np.zeros(n, F32), 1e-3, 0.9, 0.999, 1e-8, 5e-4, 1, 0.5)
        assert rel_err(down(m, ps[i], (n,)), want) < 1e-6
