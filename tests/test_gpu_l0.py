"""L0 parity on a B200: the device-module protocol of CUDA_BACKEND (-> C ABI -> sm_100a kernels) against
the oracle and the reference-derived fixtures. Copy / compare / max ops must be bit-exact."""
import ctypes
import importlib.util
import os

import numpy as np
import pytest

from conftest import golden, rel_err, ROOT

pytestmark = pytest.mark.gpu
F32 = np.float32


def dev_array(m, a):
    h = m.Array(a.size)
    m.from_numpy(np.ascontiguousarray(a, dtype=F32), h)
    return h


def host(m, h, n=None):
    n = h.size if n is None else n
    return m.to_numpy(h, (n,), (1,), 0)


def test_known_answers_of_the_reference_test(cuda_device):
    m = cuda_device.mod  # test/test_cuda.py:47-97
    a = m.Array(5)
    m.fill(a, 3.14)
    assert np.array_equal(host(m, a), np.full(5, 3.14, F32))
    x, y, o = dev_array(m, np.arange(1, 6)), dev_array(m, np.array([10, 20, 30, 40, 50])), m.Array(5)
    m.ewise_add(x, y, o)
    assert np.array_equal(host(m, o), np.array([11, 22, 33, 44, 55], F32))
    s, so = dev_array(m, np.array([1, 2, 3])), m.Array(3)
    m.scalar_add(s, 5, so)
    assert np.array_equal(host(m, so), np.array([6, 7, 8], F32))


@pytest.mark.parametrize("n", [1, 3, 4, 5, 127, 1024, 4099, 1 << 20, (1 << 22) + 3])
def test_elementwise_family_bit_exact(cuda_device, n):
    m = cuda_device.mod
    rng = np.random.RandomState(n % 1000)
    a, b = rng.randn(n).astype(F32), rng.randn(n).astype(F32)
    b[b == 0] = 1
    if n > 4:
        a[1], b[1] = 2.0, 2.0  # an exact tie for eq / ge
    ha, hb, ho = dev_array(m, a), dev_array(m, b), m.Array(n)
    for fn, want in ((m.ewise_add, a + b), (m.ewise_mul, a * b), (m.ewise_div, a / b), (m.ewise_maximum, np.maximum(a, b)),
                     (m.ewise_eq, (a == b).astype(F32)), (m.ewise_ge, (a >= b).astype(F32))):
        fn(ha, hb, ho)
        assert np.array_equal(host(m, ho), want.astype(F32)), fn.__name__
    v = F32(1.7)
    for fn, want in ((m.scalar_add, a + v), (m.scalar_mul, a * v), (m.scalar_div, a / v), (m.scalar_maximum, np.maximum(a, v)),
                     (m.scalar_eq, (a == v).astype(F32)), (m.scalar_ge, (a >= v).astype(F32))):
        fn(ha, float(v), ho)
        assert np.array_equal(host(m, ho), want.astype(F32)), fn.__name__
    m.scalar_maximum(ha, 0.0, ho)
    assert np.array_equal(host(m, ho), np.maximum(a, 0))
    # transcendental: a few ulp
    m.ewise_exp(ha, ho)
    assert rel_err(host(m, ho), np.exp(a.astype(np.float64))) < 1e-6
    m.ewise_tanh(ha, ho)
    assert np.abs(host(m, ho) - np.tanh(a.astype(np.float64))).max() < 1e-6
    m.ewise_log(ha, ho)
    got = host(m, ho)
    assert np.array_equal(got[a <= 0], np.full((a <= 0).sum(), -np.inf, F32))  # cu:405
    assert np.abs(got[a > 0] - np.log(a[a > 0].astype(np.float64))).max() < 1e-5
    pa = np.abs(a) + F32(0.1)
    m.from_numpy(pa, ha)
    m.scalar_power(ha, 0.5, ho)
    assert rel_err(host(m, ho), np.sqrt(pa.astype(np.float64))) < 1e-6
    m.scalar_power(ha, 2.0, ho)
    assert rel_err(host(m, ho), pa.astype(np.float64) ** 2) < 1e-6


def test_error_behaviour_matches_the_reference(cuda_device):
    m = cuda_device.mod
    a, b = m.Array(4), m.Array(5)
    with pytest.raises(ValueError):
        m.ewise_add(a, b, a)  # size mismatch -> std::invalid_argument
    with pytest.raises(ValueError):
        m.scalar_div(a, 0.0, a)  # cu:305 std::domain_error
    with pytest.raises(ValueError):
        m.from_numpy(np.zeros(3, F32), a)
    with pytest.raises(ValueError):
        m.compact(a, a, (1,) * 9, (1,) * 9, 0)  # > 8 dims, cu:113-118
    with pytest.raises(ValueError):
        m.fill(None, 1.0)
    m.from_numpy(np.arange(4, dtype=np.float64), a)  # forcecast f64 -> f32
    assert np.array_equal(host(m, a), np.arange(4, dtype=F32))
    assert a.size == 4 and isinstance(a.ptr(), int)


VIEWS = [  # base shape, view builder on a numpy array (must be expressible with non-negative strides)
    ((3, 4, 5, 6), lambda x: x.transpose(2, 0, 3, 1)),
    ((3, 4, 5, 6), lambda x: x[1:3, 0:4:2, 2:3, 1:6:2]),
    ((64, 48), lambda x: x.T),
    ((8, 33, 65), lambda x: x.transpose(0, 2, 1)),
    ((4, 16, 10, 10), lambda x: x.transpose(0, 2, 3, 1)),       # NCHW -> NHWC
    ((4, 10, 10, 16), lambda x: x.transpose(0, 3, 1, 2)),       # NHWC -> NCHW
    ((2, 3, 3, 3, 6, 6), lambda x: x.transpose(0, 4, 5, 1, 2, 3)),  # the reference's im2col permute, F.py:341
    ((256, 3, 32, 32), lambda x: x.transpose(0, 2, 3, 1)),
    ((7,), lambda x: x[2:6]),
]


@pytest.mark.parametrize("case", range(len(VIEWS)))
def test_compact_and_setitem_bit_exact(cuda_device, case):
    m = cuda_device.mod
    shape, view_of = VIEWS[case]
    rng = np.random.RandomState(case)
    x = rng.randn(*shape).astype(F32)
    v = view_of(x)
    strides = [s // 4 for s in v.strides]
    offset = (v.__array_interface__["data"][0] - x.__array_interface__["data"][0]) // 4
    hx = dev_array(m, x)
    out = m.Array(v.size)
    m.compact(hx, out, v.shape, strides, offset)
    assert np.array_equal(host(m, out).reshape(v.shape), v)
    assert np.array_equal(m.to_numpy(hx, v.shape, strides, offset), v)
    # scatter back into a zeroed buffer: ewise_setitem is the inverse of compact on the view
    z = m.Array(x.size)
    m.fill(z, 0.0)
    m.ewise_setitem(out, z, v.shape, strides, offset)
    want = np.zeros_like(x)
    view_of(want)[...] = v
    assert np.array_equal(host(m, z).reshape(shape), want)
    m.scalar_setitem(v.size, 2.5, z, v.shape, strides, offset)
    view_of(want)[...] = 2.5
    assert np.array_equal(host(m, z).reshape(shape), want)


REDUCE_VIEWS = [  # (shape, view with the reduced axis last)
    ((256, 2, 2, 256), lambda x: x.transpose(0, 3, 2, 1)),      # global average pool, first mean: NHWC buffer viewed (N, C, W, H)
    ((256, 256, 2), lambda x: x),                                # second mean: already compact
    ((6, 5, 7, 3), lambda x: x.transpose(1, 3, 0, 2)[1:4]),      # permuted and sliced, 7-element rows
    ((4, 32), lambda x: x[:, ::-1]),                             # negative stride along the reduced axis
    ((5, 1, 9), lambda x: np.broadcast_to(x, (5, 4, 9))),        # zero stride in the outer view
    ((33,), lambda x: x[1:33]),                                  # one output, 32 elements, offset
]


@pytest.mark.parametrize("case", range(len(REDUCE_VIEWS)))
@pytest.mark.parametrize("divisor", [1.0, 3.0, 1024.0])
def test_reduce_sum_view_div_equals_the_composed_ops(cuda_device, case, divisor):
    """dfb_reduce_sum_view_div == compact -> reduce_sum -> scalar_div (what backend_tensor.py composes for sum / mean over
    one axis) bit for bit, and == numpy's sequential float32 sum."""
    m = cuda_device.mod
    shape, view_of = REDUCE_VIEWS[case]
    rng = np.random.RandomState(100 + case)
    x = rng.randn(*shape).astype(F32)
    v = view_of(x)
    strides = [s // 4 for s in v.strides]
    offset = (v.__array_interface__["data"][0] - x.__array_interface__["data"][0]) // 4
    rows, r = int(np.prod(v.shape[:-1], dtype=np.int64)), v.shape[-1]
    hx = dev_array(m, x)
    got = m.Array(rows)
    m.reduce_sum_view_div(hx, got, v.shape, strides, offset, divisor)
    c, s1, s2 = m.Array(v.size), m.Array(rows), m.Array(rows)
    m.compact(hx, c, v.shape, strides, offset)
    m.reduce_sum(c, s1, r)
    m.scalar_div(s1, divisor, s2)
    assert np.array_equal(host(m, got), host(m, s2))
    acc = v.reshape(rows, r)[:, 0].copy()
    for j in range(1, r):
        acc = acc + v.reshape(rows, r)[:, j]
    assert np.array_equal(host(m, got), (acc / F32(divisor)).astype(F32))
    with pytest.raises(ValueError):
        m.reduce_sum_view_div(hx, got, v.shape, strides, offset, 0.0)


@pytest.mark.parametrize("case", range(len(VIEWS)))
def test_compact_scale_equals_scalar_mul_then_compact(cuda_device, case):
    m = cuda_device.mod
    shape, view_of = VIEWS[case]
    rng = np.random.RandomState(200 + case)
    x = rng.randn(*shape).astype(F32)
    v = view_of(x)
    strides = [s // 4 for s in v.strides]
    offset = (v.__array_interface__["data"][0] - x.__array_interface__["data"][0]) // 4
    hx, out, scaled, ref = dev_array(m, x), m.Array(v.size), m.Array(x.size), m.Array(v.size)
    m.compact_scale(hx, out, v.shape, strides, offset, 0.37)
    m.scalar_mul(hx, 0.37, scaled)
    m.compact(scaled, ref, v.shape, strides, offset)
    assert np.array_equal(host(m, out), host(m, ref))
    assert np.array_equal(host(m, out).reshape(v.shape), v * F32(0.37))


def test_broadcast_and_negative_strides(cuda_device):
    m = cuda_device.mod
    x = np.random.RandomState(0).randn(1, 5, 1, 7).astype(F32)
    hx = dev_array(m, x)
    out = m.Array(3 * 5 * 4 * 7)
    m.compact(hx, out, (3, 5, 4, 7), (0, 7, 0, 1), 0)
    assert np.array_equal(host(m, out).reshape(3, 5, 4, 7), np.broadcast_to(x, (3, 5, 4, 7)))
    y = np.arange(24, dtype=F32).reshape(4, 6)
    hy = dev_array(m, y)
    out = m.Array(24)
    m.compact(hy, out, (4, 6), (-6, -1), 23)  # flip both axes, bt.py:665-676
    assert np.array_equal(host(m, out).reshape(4, 6), y[::-1, ::-1])


def test_backendtensor_against_reference_fixture(cuda_device):
    from DeepFlows import backend_api
    g = golden("l0")
    x = backend_api.Btensor(g["x"], device=cuda_device)
    assert np.array_equal(x.permute((2, 0, 3, 1)).compact().numpy(), g["permute_compact"])
    assert np.array_equal(x[1:3, 0:4:2, 2, 1:6:2].compact().numpy(), g["slice_compact"])
    assert np.array_equal(x.pad(((0, 0), (0, 0), (2, 2), (1, 1))).numpy(), g["pad"])
    z = cuda_device.full((4, 6), 0.0)
    z[1:3, 0:6:2] = backend_api.Btensor(np.arange(6, dtype=F32).reshape(2, 3), device=cuda_device)
    z[3, 1:5] = 7.0
    assert np.array_equal(z.numpy(), g["setitem"])
    assert rel_err(x.sum(axis=1).numpy(), g["sum_axis1"]) < 1e-6
    assert np.array_equal(x.max(axis=2, keepdims=True).numpy(), g["max_axis2"])
    assert rel_err(x.mean(axis=2).numpy(), g["mean_axis2_quirk"]) < 1e-6
    mm = backend_api.Btensor(g["m1"], device=cuda_device) @ backend_api.Btensor(g["m2"], device=cuda_device)
    assert rel_err(mm.numpy(), g["matmul"]) < 1e-6


@pytest.mark.parametrize("rows,length", [(1, 1), (1000, 4), (257, 10), (64, 100), (33, 1025), (3, 70000), (1, 1 << 20), (1, 5)])
def test_reductions(cuda_device, rows, length):
    m = cuda_device.mod
    a = np.random.RandomState(rows).randn(rows, length).astype(F32)
    ha, ho = dev_array(m, a), m.Array(rows)
    m.reduce_max(ha, ho, length)
    assert np.array_equal(host(m, ho), a.max(axis=1))
    m.reduce_sum(ha, ho, length)
    want = a.astype(np.float64).sum(axis=1)
    assert np.abs(host(m, ho) - want).max() <= 1e-5 * np.abs(a).sum(axis=1).max()


@pytest.mark.parametrize("M,N,P", [(1, 1, 1), (7, 5, 3), (64, 64, 64), (100, 300, 50), (256, 784, 100), (33, 2049, 65), (1024, 64, 10)])
def test_matmul_fp32(cuda_device, M, N, P):
    m = cuda_device.mod
    rng = np.random.RandomState(M + N + P)
    a, b = rng.randn(M, N).astype(F32), rng.randn(N, P).astype(F32)
    out = m.Array(M * P)
    m.set_matmul_mode(m.MODE_FP32)
    m.matmul(dev_array(m, a), dev_array(m, b), out, M, N, P)
    want = a.astype(np.float64) @ b.astype(np.float64)
    assert rel_err(host(m, out).reshape(M, P), want) < 1e-5  # north_star: 1e-5 relative in fp32 mode


def test_c_abi_direct(cuda_device):
    """The same path without Python objects: raw pointers through the extern "C" entry points."""
    import deepflows_b200
    lib = ctypes.CDLL(deepflows_b200.lib_path())
    n = 1000
    a = np.random.RandomState(0).randn(n).astype(F32)
    b = np.random.RandomState(1).randn(n).astype(F32)
    out = np.empty(n, F32)
    pa, pb, po = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
    for p in (pa, pb, po):
        assert lib.dfb_malloc(ctypes.c_size_t(n), ctypes.byref(p)) == 0
    fptr = ctypes.POINTER(ctypes.c_float)
    assert lib.dfb_from_host(a.ctypes.data_as(fptr), pa, ctypes.c_size_t(n)) == 0
    assert lib.dfb_from_host(b.ctypes.data_as(fptr), pb, ctypes.c_size_t(n)) == 0
    assert lib.dfb_ewise_mul(pa, pb, po, ctypes.c_size_t(n)) == 0
    assert lib.dfb_to_host(po, out.ctypes.data_as(fptr), ctypes.c_size_t(n)) == 0
    assert np.array_equal(out, a * b)
    lib.dfb_launch_count.restype = ctypes.c_uint64
    assert lib.dfb_launch_count() > 0
    assert lib.dfb_scalar_div(pa, ctypes.c_float(0.0), po, ctypes.c_size_t(n)) == 5  # DFB_ERR_DOMAIN
    for p in (pa, pb, po):
        assert lib.dfb_free(p) == 0


def _reference_cuda_module():
    path = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.isdir(path):
        return None
    for f in os.listdir(path):
        if f.startswith("CUDA_BACKEND") and f.endswith(".so"):
            spec = importlib.util.spec_from_file_location("CUDA_BACKEND", os.path.join(path, f))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            return mod
    return None


def test_against_the_reference_cuda_module(cuda_device):
    """oracle/_ref holds the reference's own ndarray_backend_cuda.cu compiled for sm_100 (oracle/Makefile).
    Same inputs through both modules; copy/compare/arith ops bit-exact, matmul/reduce within 1e-5."""
    ref = _reference_cuda_module()
    if ref is None:
        pytest.skip("oracle/_ref not built")
    m = cuda_device.mod
    rng = np.random.RandomState(5)
    n = 4096 + 3
    a, b = rng.randn(n).astype(F32), rng.randn(n).astype(F32)

    def both(fn_name, *extra):
        outs = []
        for mod in (m, ref):
            ha, hb, ho = mod.Array(n), mod.Array(n), mod.Array(n)
            mod.from_numpy(a, ha)
            mod.from_numpy(b, hb)
            fn = getattr(mod, fn_name)
            if fn_name.startswith("ewise_") and fn_name not in ("ewise_log", "ewise_exp", "ewise_tanh"):
                fn(ha, hb, ho)
            elif fn_name.startswith("scalar_"):
                fn(ha, *extra, ho)
            else:
                fn(ha, ho)
            outs.append(mod.to_numpy(ho, (n,), (1,), 0))
        return outs

    for name in ("ewise_add", "ewise_mul", "ewise_div", "ewise_maximum", "ewise_eq", "ewise_ge"):
        mine, theirs = both(name)
        assert np.array_equal(mine, theirs), name
    for name, v in (("scalar_add", 1.5), ("scalar_mul", -2.0), ("scalar_div", 3.0), ("scalar_maximum", 0.0),
                    ("scalar_eq", 0.0), ("scalar_ge", 0.25)):
        mine, theirs = both(name, v)
        assert np.array_equal(mine, theirs), name
    for name in ("ewise_exp", "ewise_tanh", "ewise_log"):
        mine, theirs = both(name)
        fin = np.isfinite(theirs)
        assert np.array_equal(np.isfinite(mine), fin), name
        assert np.abs(mine[fin] - theirs[fin]).max() <= 4e-7 * max(1.0, np.abs(theirs[fin]).max()), name
    # strided gather / scatter
    x = rng.randn(4, 6, 5, 7).astype(F32)
    shape, strides = (5, 4, 7, 6), (7, 210, 1, 35)
    outs = []
    for mod in (m, ref):
        hx, ho, hz = mod.Array(x.size), mod.Array(x.size), mod.Array(x.size)
        mod.from_numpy(x, hx)
        mod.compact(hx, ho, shape, strides, 0)
        mod.fill(hz, 0.0)
        hs = mod.Array(420)  # exactly the view's element count, as BackendTensor.__setitem__ guarantees (bt.py:512)
        mod.from_numpy(x.reshape(-1)[:420].copy(), hs)
        mod.ewise_setitem(hs, hz, (4, 3, 5, 7), (210, 70, 7, 1), 35)
        mod.scalar_setitem(10, 9.0, hz, (2, 5), (7, 1), 0)
        outs.append((mod.to_numpy(ho, (x.size,), (1,), 0), mod.to_numpy(hz, (x.size,), (1,), 0)))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    # matmul and reductions
    M, N, P = 65, 130, 33
    A, B = rng.randn(M, N).astype(F32), rng.randn(N, P).astype(F32)
    res = []
    for mod in (m, ref):
        hA, hB, hC = mod.Array(M * N), mod.Array(N * P), mod.Array(M * P)
        mod.from_numpy(A, hA)
        mod.from_numpy(B, hB)
        mod.matmul(hA, hB, hC, M, N, P)
        hs, hm = mod.Array(M), mod.Array(M)
        mod.reduce_sum(hA, hs, N)
        mod.reduce_max(hA, hm, N)
        res.append((mod.to_numpy(hC, (M, P), (P, 1), 0), mod.to_numpy(hs, (M,), (1,), 0), mod.to_numpy(hm, (M,), (1,), 0)))
    assert rel_err(res[0][0], res[1][0]) < 1e-5 and rel_err(res[0][1], res[1][1]) < 1e-5
    assert np.array_equal(res[0][2], res[1][2])
