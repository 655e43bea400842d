"""SURVEY 8f rank 4 ops against plain numpy restatements: general pooling (stride != kernel, zero padding), the 1-d
ops, the activation modules, Softmax, and device-state Adagrad / Adadelta. Shared by the CPU tier (oracle numpy device) and
the GPU tier."""
import numpy as np

from conftest import rel_err

F32 = np.float32


def _windows2d(x, k, s, p):
    xp = np.pad(x, ((0, 0), (0, 0), (p, p), (p, p)))
    n, c, h, w = xp.shape
    oh, ow = (h - k) // s + 1, (w - k) // s + 1
    col = np.zeros((n, c, oh, ow, k, k), x.dtype)
    for i in range(k):
        for j in range(k):
            col[:, :, :, :, i, j] = xp[:, :, i:i + oh * s:s, j:j + ow * s:s]
    return col, xp.shape


def check_general_pooling(device):
    from DeepFlows import backend_api, tensor
    from DeepFlows.nn import functional as F
    from DeepFlows.tensor import Tensor
    rng = np.random.RandomState(0)
    backend_api.set_dgrad_mode("exact")
    for (k, s, p, shape) in [(3, 2, 1, (2, 3, 9, 9)), (2, 1, 0, (1, 4, 5, 6)), (3, 1, 1, (2, 2, 6, 6)), (3, 3, 1, (1, 2, 7, 7))]:
        for is_max in (True, False):
            tensor.Graph.free_graph_all()
            x = np.round(rng.randn(*shape) * 2).astype(F32) / 2 if is_max else rng.randn(*shape).astype(F32)   # ties for max
            xt = Tensor(x, device=device, requires_grad=True)
            y = (F.max_pool2d if is_max else F.avg_pool2d)(xt, k, s, p)
            col, pshape = _windows2d(x, k, s, p)
            want = col.max(axis=(4, 5)) if is_max else col.mean(axis=(4, 5))
            yv = y.numpy()
            assert np.array_equal(yv, want) if is_max else rel_err(yv, want) < 1e-6, (k, s, p, is_max)
            gy = rng.randn(*yv.shape).astype(F32)
            tensor.sum(y * Tensor(gy, device=device)).backward()
            # textbook gradient with the reference's tie rule: every element equal to its window's maximum receives it
            gp = np.zeros(pshape, np.float64)
            n, c, oh, ow = want.shape
            for i in range(k):
                for j in range(k):
                    m = (col[:, :, :, :, i, j] == want) if is_max else np.full(want.shape, 1.0 / (k * k))
                    gp[:, :, i:i + oh * s:s, j:j + ow * s:s] += m * gy
            gx = gp[:, :, p:gp.shape[2] - p, p:gp.shape[3] - p]
            assert rel_err(xt.grad.numpy(), gx) < 1e-6, (k, s, p, is_max)


def check_1d_ops(device):
    from DeepFlows import tensor
    from DeepFlows.nn import functional as F
    from DeepFlows.tensor import Tensor
    rng = np.random.RandomState(1)
    for (n, c, l, ko, k, s, p) in [(2, 3, 11, 4, 3, 1, 1), (1, 2, 10, 5, 4, 2, 0), (3, 1, 7, 2, 2, 1, 2)]:
        tensor.Graph.free_graph_all()
        x, w = rng.randn(n, c, l).astype(F32), rng.randn(ko, c, k).astype(F32)
        xt, wt = Tensor(x, device=device, requires_grad=True), Tensor(w, device=device, requires_grad=True)
        y = F.conv1d(xt, wt, p, s)
        xp = np.pad(x, ((0, 0), (0, 0), (p, p))).astype(np.float64)
        lo = (xp.shape[2] - k) // s + 1
        col = np.stack([xp[:, :, i:i + lo * s:s] for i in range(k)], axis=2)          # (N, C, k, Lo)
        want = np.einsum("nckl,ock->nol", col, w.astype(np.float64))
        assert y.shape == (n, ko, lo) and rel_err(y.numpy(), want) < 1e-5
        gy = rng.randn(n, ko, lo).astype(F32)
        tensor.sum(y * Tensor(gy, device=device)).backward()
        dw = np.einsum("nol,nckl->ock", gy.astype(np.float64), col)
        dcol = np.einsum("nol,ock->nckl", gy.astype(np.float64), w.astype(np.float64))
        dxp = np.zeros_like(xp)
        for i in range(k):
            dxp[:, :, i:i + lo * s:s] += dcol[:, :, i]
        assert rel_err(wt.grad.numpy(), dw) < 1e-5
        assert rel_err(xt.grad.numpy(), dxp[:, :, p:dxp.shape[2] - p]) < 1e-5
        for fn, red in ((F.max_pool1d, lambda a: a.max(axis=2)), (F.avg_pool1d, lambda a: a.mean(axis=2))):
            tensor.Graph.free_graph_all()
            out = fn(Tensor(x, device=device), k, s, p)
            assert rel_err(out.numpy(), red(col)) < 1e-6


def check_activations(device):
    from DeepFlows import nn, tensor
    from DeepFlows.tensor import Tensor
    rng = np.random.RandomState(2)
    x = (rng.randn(6, 10) * 3).astype(F32)
    gy = rng.randn(6, 10).astype(F32)
    x64 = x.astype(np.float64)
    sm = np.exp(x64 - x64.max(1, keepdims=True))
    sm /= sm.sum(1, keepdims=True)
    cases = {
        "Sigmoid": (nn.Sigmoid(), 1 / (1 + np.exp(-x64)), lambda y: y * (1 - y) * gy),
        "Tanh": (nn.Tanh(), np.tanh(x64), lambda y: (1 - y * y) * gy),
        "LeakyReLU": (nn.LeakyReLU(0.1), np.where(x64 > 0, x64, 0.1 * x64), lambda y: np.where(x64 > 0, 1.0, 0.1) * gy),
        "Softmax": (nn.Softmax(dim=1), sm, lambda y: y * (gy - (gy * y).sum(1, keepdims=True))),
    }
    for name, (mod, want, grad) in cases.items():
        tensor.Graph.free_graph_all()
        xt = Tensor(x, device=device, requires_grad=True)
        y = mod(xt)
        assert rel_err(y.numpy(), want) < 2e-6, name
        tensor.sum(y * Tensor(gy, device=device)).backward()
        assert rel_err(xt.grad.numpy(), grad(want)) < 2e-5, name


def check_adagrad_adadelta(device):
    from DeepFlows import backend_api, tensor
    from DeepFlows.optim import Adagrad, Adadelta
    from DeepFlows.tensor import Tensor
    rng = np.random.RandomState(3)
    shapes = [(5, 4), (1, 4), (2, 3, 3, 3)]
    for name in ("adagrad", "adadelta"):
        p0 = [rng.randn(*s).astype(F32) for s in shapes]
        grads = [[rng.randn(*s).astype(F32) for s in shapes] for _ in range(3)]
        tensor.Graph.free_graph_all()
        params = [Tensor(p, device=device, requires_grad=True) for p in p0]
        opt = Adagrad(params, lr=0.05, weight_decay=1e-3) if name == "adagrad" else Adadelta(params, rho=0.9, weight_decay=1e-3)
        ref = [p.astype(np.float64) for p in p0]
        s_acc = [np.zeros_like(p) for p in ref]
        d_acc = [np.zeros_like(p) for p in ref]
        for gs in grads:
            for p, g in zip(params, gs):
                p.grad = backend_api.Btensor(g, device=device)
            opt.step()
            for i, g in enumerate(gs):          # reference: optim/adagrad.py, optim/adadelta.py (numpy state there)
                g = g.astype(np.float64) + ref[i] * 1e-3
                if name == "adagrad":
                    s_acc[i] += g * g
                    ref[i] = ref[i] - 0.05 * g / np.sqrt(s_acc[i] + 1e-10)
                else:
                    s_acc[i] = 0.9 * s_acc[i] + 0.1 * g * g
                    upd = g * np.sqrt(d_acc[i] + 1e-6) / np.sqrt(s_acc[i] + 1e-6)
                    d_acc[i] = 0.9 * d_acc[i] + 0.1 * upd * upd
                    ref[i] = ref[i] - upd
        for p, r in zip(params, ref):
            assert rel_err(p.data.numpy(), r) < 1e-5, name


def check_device_dropout(device, mask_of):
    """Opt-in device RNG for Dropout: the Philox mask kernel against its oracle restatement (bit-exact), the published
    Philox4x32-10 known answer, and the module path (train scaling 1/(1-p), a new mask every step, reproducible by seed)."""
    from oracle import numpy_ops as ops
    from DeepFlows import backend_api, nn, tensor
    from DeepFlows.tensor import Tensor
    # Random123 known-answer test: counter (0,0,0,0), key (0,0) -> 6627e8d5 e169c58d bc57ac4c 9b00dbd8
    g = np.zeros(1, np.uint64)
    c = [g.copy(), g.copy(), g.copy(), g.copy()]
    k0, k1, m32 = np.uint64(0), np.uint64(0), np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = np.uint64(0xD2511F53) * c[0], np.uint64(0xCD9E8D57) * c[2]
        c = [(p1 >> np.uint64(32)) ^ c[1] ^ k0, p1 & m32, (p0 >> np.uint64(32)) ^ c[3] ^ k1, p0 & m32]
        k0, k1 = (k0 + np.uint64(0x9E3779B9)) & m32, (k1 + np.uint64(0xBB67AE85)) & m32
    assert [int(v[0]) for v in c] == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    for n, keep, seed, step in [(1000, 0.5, 12345, 1), (4099, 0.8, 7, 300), (8, 0.25, 16777215, 2)]:
        assert np.array_equal(mask_of(n, keep, seed, step), ops.philox_dropout_mask(n, keep, seed, step)), (n, keep, seed, step)
    backend_api.set_dropout_rng("device")
    try:
        outs = []
        for _ in range(2):
            tensor.Graph.free_graph_all()
            np.random.seed(3)
            drop = nn.Dropout(0.25)
            drop.train()
            x = Tensor(np.ones((64, 128), F32), device=device)
            a, b = drop(x).numpy(), drop(x).numpy()
            assert set(np.unique(a)) <= {0.0, np.float32(1 / 0.75)} and 0.6 < (a > 0).mean() < 0.9
            assert not np.array_equal(a, b), "the second step must draw a new mask"
            outs.append((a, b))
        assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1]), "same numpy seed, same masks"
    finally:
        backend_api.set_dropout_rng("host")
