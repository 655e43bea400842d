"""The oracle (oracle/numpy_ops.py) against the committed fixtures that oracle/make_golden.py took
from the reference itself, and against the reference's own known answers (test/test_cuda.py:47-97)."""
import numpy as np
import pytest

from conftest import golden, rel_err
from oracle import numpy_ops as ops
from oracle import numpy_device as nd

F32 = np.float32


def test_reference_known_answers():
    # test/test_cuda.py:47-97 — the only known-answer checks the reference holds
    a = nd.Array(5)
    nd.fill(a, 3.14)
    assert np.array_equal(a.buf, np.full(5, 3.14, F32))
    x, y, o = nd.Array(5), nd.Array(5), nd.Array(5)
    nd.from_numpy(np.arange(1, 6, dtype=F32), x)
    nd.from_numpy(np.array([10, 20, 30, 40, 50], F32), y)
    nd.ewise_add(x, y, o)
    assert np.array_equal(o.buf, np.array([11, 22, 33, 44, 55], F32))
    s, so = nd.Array(3), nd.Array(3)
    nd.from_numpy(np.array([1, 2, 3], F32), s)
    nd.scalar_add(s, 5, so)
    assert np.array_equal(so.buf, np.array([6, 7, 8], F32))


def test_strided_ops_match_reference():
    g = golden("l0")
    x = g["x"]
    flat = x.reshape(-1)
    st = np.array(x.strides) // 4
    perm = (2, 0, 3, 1)
    got = ops.compact(flat, [x.shape[p] for p in perm], [st[p] for p in perm], 0)
    assert np.array_equal(got.reshape(g["permute_compact"].shape), g["permute_compact"])
    out = np.zeros(24, F32)
    ops.ewise_setitem(np.arange(6, dtype=F32), out, (2, 3), (6, 2), 6)
    ops.scalar_setitem(4, 7.0, out, (1, 4), (6, 1), 19)
    assert np.array_equal(out.reshape(4, 6), g["setitem"])


def test_log_nonpositive_is_minus_inf():
    # ndarray_backend_cuda.cu:405
    assert np.array_equal(ops.ewise_log(np.array([-1.0, 0.0, 1.0], F32)), np.array([-np.inf, -np.inf, 0.0], F32))


@pytest.mark.parametrize("case", ["k3p1s1", "k5p2s1", "k3p1s2", "k1p0s2", "stem_c3", "k3p0s1_rect", "k3p1s2_odd"])
def test_conv_matches_reference(case):
    g = golden("conv")
    n, c, h, w, k, r, p, s = g[case + ".geom"]
    x, wt, gy = g[case + ".x"], g[case + ".w"], g[case + ".gy"]
    assert rel_err(ops.conv2d_fprop(x, wt, p, s), g[case + ".y"]) < 2e-5
    assert rel_err(ops.conv2d_dgrad_reference(gy, wt, x.shape, p, s), g[case + ".dx_ref"]) < 2e-5
    assert rel_err(ops.conv2d_wgrad(x, gy, wt.shape, p, s), g[case + ".dw"]) < 2e-5


def test_dgrad_exact_equals_reference_only_without_overlap():
    g = golden("conv")
    # 1x1 stride 2: windows do not overlap, last-writer-wins == true gradient
    gy, wt, x = g["k1p0s2.gy"], g["k1p0s2.w"], g["k1p0s2.x"]
    assert rel_err(ops.conv2d_dgrad_exact(gy, wt, x.shape, 0, 2), g["k1p0s2.dx_ref"]) < 2e-5
    # 3x3 stride 1: they differ
    gy, wt, x = g["k3p1s1.gy"], g["k3p1s1.w"], g["k3p1s1.x"]
    assert rel_err(ops.conv2d_dgrad_exact(gy, wt, x.shape, 1, 1), g["k3p1s1.dx_ref"]) > 1e-2


def test_dgrad_exact_is_adjoint_of_fprop():
    rng = np.random.RandomState(0)
    for (n, c, h, w, k, r, p, s) in [(2, 3, 9, 9, 4, 3, 1, 2), (1, 4, 8, 8, 5, 5, 2, 1), (2, 2, 7, 7, 3, 1, 0, 2)]:
        x = rng.randn(n, c, h, w).astype(F32)
        wt = rng.randn(k, c, r, r).astype(F32)
        y = ops.conv2d_fprop(x, wt, p, s)
        gy = rng.randn(*y.shape).astype(F32)
        lhs = float((y.astype(np.float64) * gy).sum())
        rhs = float((ops.conv2d_dgrad_exact(gy, wt, x.shape, p, s).astype(np.float64) * x).sum())
        rhs_w = float((ops.conv2d_wgrad(x, gy, wt.shape, p, s).astype(np.float64) * wt).sum())
        assert abs(lhs - rhs) < 1e-3 * max(1.0, abs(lhs))
        assert abs(lhs - rhs_w) < 1e-3 * max(1.0, abs(lhs))


def test_bn_pool_relu_ce_match_reference():
    g = golden("ops")
    c = g["bn.x"].shape[1]
    y, rm, rv, _, _ = ops.bn_fwd_train(g["bn.x"], g["bn.gamma"], g["bn.beta"], np.zeros((1, c, 1, 1), F32),
                                       np.ones((1, c, 1, 1), F32), 0.1, 1e-5)
    assert rel_err(y, g["bn.y"]) < 2e-5 and rel_err(rm, g["bn.running_mean"]) < 2e-5 and rel_err(rv, g["bn.running_var"]) < 2e-5
    dx, dg, db = ops.bn_bwd(g["bn.x"], g["bn.gy"], g["bn.gamma"], 1e-5)
    assert rel_err(dx, g["bn.dx"]) < 5e-5 and rel_err(dg, g["bn.dgamma"]) < 5e-5 and rel_err(db, g["bn.dbeta"]) < 5e-5
    assert rel_err(ops.bn_fwd_eval(g["bn.x"], g["bn.gamma"], g["bn.beta"], g["bn.running_mean"], g["bn.running_var"], 1e-5),
                   g["bn.y_eval"]) < 2e-5
    assert np.array_equal(ops.relu_fwd(g["relu.x"]), g["relu.y"])
    assert np.array_equal(ops.relu_bwd(g["relu.x"], g["relu.gy"]), g["relu.dx"])
    for name in ("pool", "pool_odd"):
        assert np.array_equal(ops.maxpool2d_fwd(g[name + ".x"], 2), g[name + ".y"])
        assert np.array_equal(ops.maxpool2d_bwd(g[name + ".x"], g[name + ".y"], g[name + ".gy"], 2), g[name + ".dx"])
    for name, scale in (("ce_mean", 1 / 16), ("ce_sum", 1.0)):
        assert rel_err(ops.softmax_ce_fwd(g[name + ".logits"], g[name + ".target"], scale), g[name + ".loss"]) < 2e-5
        assert rel_err(ops.softmax_ce_bwd(g[name + ".logits"], g[name + ".target"], 1.0, scale), g[name + ".dlogits"]) < 5e-5


def test_optimizers_match_reference():
    g = golden("optim")
    cfg = {"adam": dict(lr=5e-3, wd=5e-4), "adam_nowd": dict(lr=1e-3, wd=0.0)}
    for name, c in cfg.items():
        for i in range(3):
            p = g["%s.p0.%d" % (name, i)]
            v, s = np.zeros_like(p), np.zeros_like(p)
            for st in range(3):
                p, v, s = ops.adam_step(p, g["%s.g%d.%d" % (name, st, i)], v, s, c["lr"], 0.9, 0.999, 1e-8, c["wd"], st + 1)
            assert rel_err(p, g["%s.p3.%d" % (name, i)]) < 1e-6
    for name, c in {"sgd": dict(mom=0.0, wd=0.0, nes=False), "sgd_mom": dict(mom=0.9, wd=1e-3, nes=True)}.items():
        for i in range(3):
            p = g["%s.p0.%d" % (name, i)]
            v = np.zeros_like(p)
            for st in range(3):
                p, nv = ops.sgd_step(p, g["%s.g%d.%d" % (name, st, i)], v, 0.05, c["mom"], c["wd"], c["nes"])
                v = nv if nv is not None else v
            assert rel_err(p, g["%s.p3.%d" % (name, i)]) < 1e-6
