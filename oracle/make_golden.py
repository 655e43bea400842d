"""Generates tests/golden/*.npz from the REFERENCE ITSELF and pins oracle/numpy_ops.py against it.

Runs only in the build container (it imports the reference package from /root/reference, which does
not exist on the GPU box); the fixtures it writes are committed and travel. Usage:

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py

What "the reference" means here: the reference's own host code (tensor.py, nn/, optim/)
running on its numpy device, with that device's two setitem functions replaced by a restatement of the
CUDA kernels (ndarray_backend_cuda.cu:147-155,178-221: out[offset + sum idx_d*stride_d] = a[gid]),
because the as-shipped numpy setitem writes to the front of the buffer (backend_tensor.py:92-100).
That is what `device='cuda'` computes in the reference.

For every op the script (1) runs the reference, (2) runs oracle/numpy_ops.py on the same inputs and
asserts agreement (bit-exact for copies / compares / max, 2e-5 relative otherwise), (3) stores inputs and
reference outputs.
"""
import os
import sys

os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("DEEPFLOWS_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.path.insert(1, ROOT)

import numpy as np  # noqa: E402

import DeepFlows  # noqa: E402  (the reference package)
from DeepFlows.backend import backend_tensor as rbt  # noqa: E402
from DeepFlows import tensor as rtensor  # noqa: E402
from DeepFlows.tensor import Tensor  # noqa: E402
from DeepFlows import nn as rnn  # noqa: E402
from DeepFlows.nn import functional as RF  # noqa: E402
from DeepFlows.optim import Adam, SGD  # noqa: E402

from oracle import numpy_ops as ops  # noqa: E402
import workloads  # noqa: E402

assert os.path.realpath(DeepFlows.__file__).startswith(os.path.realpath(REF)), DeepFlows.__file__
GOLD = os.path.join(ROOT, "tests", "golden")
os.makedirs(GOLD, exist_ok=True)

# ---- the 2-function restatement of the CUDA setitem kernels on the reference's numpy device --------------
_orig_cpu_numpy = rbt.cpu_numpy


def _cuda_semantics_numpy_device():
    dev = _orig_cpu_numpy()
    mod = dev.mod

    def ewise_setitem(a, out, shape, strides, offset):
        ops.ewise_setitem(a, out, shape, strides, offset)

    def scalar_setitem(size, value, out, shape, strides, offset):
        ops.scalar_setitem(size, value, out, shape, strides, offset)

    mod.ewise_setitem = ewise_setitem
    mod.scalar_setitem = scalar_setitem
    return dev


rbt.cpu_numpy = _cuda_semantics_numpy_device
CPU = rbt.Device("cpu")
F32 = np.float32


def T(a, grad=False):
    return Tensor(np.asarray(a, dtype=F32), device=CPU, requires_grad=grad)


def close(name, got, want, tol=2e-5):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, (name, got.shape, want.shape)
    scale = max(np.abs(want).max(), 1e-30)
    err = np.abs(got - want).max() / scale
    assert err <= tol, "%s: oracle differs from the reference by %.3g (relative to max |ref|)" % (name, err)
    return err


def exact(name, got, want):
    assert np.array_equal(np.asarray(got), np.asarray(want)), "%s: not bit-exact" % name


def fresh():
    rtensor.Graph.free_graph_all()


# ------------------------------------------------------------------------------------------------
def golden_l0():
    """Known answers of the reference's own test (test/test_cuda.py:47-97) + strided ops."""
    out = {}
    a = rbt.BackendTensor(np.arange(1, 6, dtype=F32), device=CPU)
    b = rbt.BackendTensor(np.array([10, 20, 30, 40, 50], dtype=F32), device=CPU)
    f = CPU.full((5,), 3.14)
    exact("fill", f.numpy(), np.full(5, 3.14, dtype=F32))
    exact("ewise_add", (a + b).numpy(), np.array([11, 22, 33, 44, 55], dtype=F32))
    exact("scalar_add", (rbt.BackendTensor(np.array([1, 2, 3], dtype=F32), device=CPU) + 5).numpy(),
          np.array([6, 7, 8], dtype=F32))
    rng = np.random.RandomState(1)
    x = rng.randn(3, 4, 5, 6).astype(F32)
    xt = rbt.BackendTensor(x, device=CPU)
    out["x"] = x
    out["permute_compact"] = xt.permute((2, 0, 3, 1)).compact().numpy()
    exact("compact", out["permute_compact"], x.transpose(2, 0, 3, 1))
    out["slice_compact"] = xt[1:3, 0:4:2, 2, 1:6:2].compact().numpy()
    exact("slice", out["slice_compact"], x[1:3, 0:4:2, 2:3, 1:6:2])
    out["bcast"] = rbt.BackendTensor(x[:1, :, :1, :], device=CPU).broadcast_to((3, 4, 5, 6)).compact().numpy()
    pad = xt.pad(((0, 0), (0, 0), (2, 2), (1, 1)))
    out["pad"] = pad.numpy()
    exact("pad", out["pad"], np.pad(x, ((0, 0), (0, 0), (2, 2), (1, 1))))
    z = CPU.full((4, 6), 0.0)
    z[1:3, 0:6:2] = rbt.BackendTensor(np.arange(6, dtype=F32).reshape(2, 3), device=CPU)
    z[3, 1:5] = 7.0
    out["setitem"] = z.numpy()
    want = np.zeros((4, 6), dtype=F32)
    want[1:3, 0:6:2] = np.arange(6).reshape(2, 3)
    want[3, 1:5] = 7
    exact("setitem", out["setitem"], want)
    out["sum_axis1"] = xt.sum(axis=1).numpy()
    out["max_axis2"] = xt.max(axis=2, keepdims=True).numpy()
    out["mean_axis2_quirk"] = xt.mean(axis=2).numpy()  # divides by x.size, not by the axis length (Q3)
    close("mean quirk", out["mean_axis2_quirk"], x.sum(axis=2) / x.size)
    m1, m2 = rng.randn(7, 5).astype(F32), rng.randn(5, 3).astype(F32)
    out["m1"], out["m2"] = m1, m2
    out["matmul"] = (rbt.BackendTensor(m1, device=CPU) @ rbt.BackendTensor(m2, device=CPU)).numpy()
    np.savez_compressed(os.path.join(GOLD, "l0.npz"), **out)


CONV_CASES = [  # name, N, C, H, W, K, R, pad, stride
    ("k3p1s1", 2, 4, 8, 8, 8, 3, 1, 1),
    ("k5p2s1", 2, 3, 9, 9, 4, 5, 2, 1),
    ("k3p1s2", 2, 8, 8, 8, 16, 3, 1, 2),
    ("k1p0s2", 2, 8, 8, 8, 16, 1, 0, 2),
    ("stem_c3", 2, 3, 12, 12, 8, 3, 1, 1),
    ("k3p0s1_rect", 1, 5, 7, 10, 6, 3, 0, 1),
    ("k3p1s2_odd", 2, 4, 9, 9, 4, 3, 1, 2),
]


def golden_conv():
    out = {}
    rng = np.random.RandomState(2)
    for name, n, c, h, w, k, r, p, s in CONV_CASES:
        fresh()
        x = rng.randn(n, c, h, w).astype(F32)
        wt = (rng.randn(k, c, r, r) * 0.2).astype(F32)
        xt, wtt = T(x, True), T(wt, True)
        y = RF.conv2d(xt, wtt, p, s)
        yv = y.numpy()
        gy = rng.randn(*yv.shape).astype(F32)
        loss = rtensor.sum(y * T(gy))
        loss.backward()
        dx, dw = xt.grad.numpy(), wtt.grad.numpy()
        close(name + " fprop", ops.conv2d_fprop(x, wt, p, s), yv)
        close(name + " dgrad(reference)", ops.conv2d_dgrad_reference(gy, wt, x.shape, p, s), dx)
        close(name + " wgrad", ops.conv2d_wgrad(x, gy, wt.shape, p, s), dw)
        for key, val in (("x", x), ("w", wt), ("y", yv), ("gy", gy), ("dx_ref", dx), ("dw", dw)):
            out[name + "." + key] = val
        out[name + ".geom"] = np.array([n, c, h, w, k, r, p, s])
    np.savez_compressed(os.path.join(GOLD, "conv.npz"), **out)


def golden_bn_pool_act_loss():
    out = {}
    rng = np.random.RandomState(3)
    # BatchNorm2d: two training steps (running stats), gradients of x / gamma / beta, then eval
    fresh()
    n, c, h, w = 4, 6, 5, 5
    bn = rnn.BatchNorm2d(c, device="cpu")
    g0, b0 = rng.rand(1, c, 1, 1).astype(F32) + 0.5, rng.randn(1, c, 1, 1).astype(F32)
    bn.weight.data = rbt.BackendTensor(g0, device=CPU)
    bn.bias.data = rbt.BackendTensor(b0, device=CPU)
    x = (rng.randn(n, c, h, w) * 2 + 3).astype(F32)
    gy = rng.randn(n, c, h, w).astype(F32)
    bn.train()
    xt = T(x, True)
    y = bn(xt)
    rtensor.sum(y * T(gy)).backward()
    out.update({"bn.x": x, "bn.gamma": g0, "bn.beta": b0, "bn.gy": gy, "bn.y": y.numpy(), "bn.dx": xt.grad.numpy(),
                "bn.dgamma": bn.weight.grad.numpy(), "bn.dbeta": bn.bias.grad.numpy(),
                "bn.running_mean": bn.running_mean.numpy(), "bn.running_var": bn.running_var.numpy()})
    oy, nrm, nrv, _, _ = ops.bn_fwd_train(x, g0, b0, np.zeros((1, c, 1, 1), F32), np.ones((1, c, 1, 1), F32), 0.1, 1e-5)
    close("bn fwd", oy, out["bn.y"])
    close("bn running_mean", nrm, out["bn.running_mean"])
    close("bn running_var", nrv, out["bn.running_var"])
    odx, odg, odb = ops.bn_bwd(x, gy, g0, 1e-5)
    close("bn dx", odx, out["bn.dx"], 5e-5)
    close("bn dgamma", odg, out["bn.dgamma"], 5e-5)
    close("bn dbeta", odb, out["bn.dbeta"], 5e-5)
    bn.eval()
    out["bn.y_eval"] = bn(T(x)).numpy()
    close("bn eval", ops.bn_fwd_eval(x, g0, b0, out["bn.running_mean"], out["bn.running_var"], 1e-5), out["bn.y_eval"])
    bn.train()

    # ReLU incl. exact zeros (gradient passes at x == 0)
    fresh()
    x = rng.randn(3, 4, 4, 4).astype(F32)
    x[0, 0, 0, :2] = 0.0
    gy = rng.randn(*x.shape).astype(F32)
    xt = T(x, True)
    y = RF.relu(xt)
    rtensor.sum(y * T(gy)).backward()
    out.update({"relu.x": x, "relu.gy": gy, "relu.y": y.numpy(), "relu.dx": xt.grad.numpy()})
    exact("relu fwd", ops.relu_fwd(x), out["relu.y"])
    exact("relu bwd", ops.relu_bwd(x, gy), out["relu.dx"])

    # MaxPool 2x2/2 with ties (post-ReLU zeros) and an odd size (last row/col dropped)
    for name, shape in (("pool", (2, 3, 8, 8)), ("pool_odd", (2, 3, 7, 9))):
        fresh()
        x = np.maximum(rng.randn(*shape), 0).astype(F32)
        xt = T(x, True)
        y = RF.max_pool2d(xt, 2, 2)
        yv = y.numpy()
        gy = rng.randn(*yv.shape).astype(F32)
        rtensor.sum(y * T(gy)).backward()
        out.update({name + ".x": x, name + ".gy": gy, name + ".y": yv, name + ".dx": xt.grad.numpy()})
        exact(name + " fwd", ops.maxpool2d_fwd(x, 2), yv)
        exact(name + " bwd", ops.maxpool2d_bwd(x, yv, gy, 2), out[name + ".dx"])
        out[name + ".argmax"] = ops.maxpool2d_argmax(x, 2)

    # cross entropy: one-hot and label-smoothed targets, mean and sum
    for name, red in (("ce_mean", "mean"), ("ce_sum", "sum")):
        fresh()
        logits = (rng.randn(16, 10) * 3).astype(F32)
        tgt = np.eye(10, dtype=F32)[rng.randint(0, 10, 16)]
        if red == "mean":
            tgt = (tgt * (1 - 0.05) + 0.05 / 10).astype(F32)
        lt = T(logits, True)
        loss = RF.cross_entropy(lt, T(tgt), reduction=red)
        loss.backward()
        out.update({name + ".logits": logits, name + ".target": tgt, name + ".loss": loss.numpy(),
                    name + ".dlogits": lt.grad.numpy()})
        scale = 1.0 / 16 if red == "mean" else 1.0
        close(name + " fwd", ops.softmax_ce_fwd(logits, tgt, scale), out[name + ".loss"])
        close(name + " bwd", ops.softmax_ce_bwd(logits, tgt, 1.0, scale), out[name + ".dlogits"], 5e-5)

    # global average pooling as the ResNet script does it: mean(axis=2) twice, with quirk Q3
    fresh()
    x = rng.randn(3, 5, 2, 2).astype(F32)
    xt = T(x, True)
    y = rtensor.mean(rtensor.mean(xt, axis=2), axis=2)
    gy = rng.randn(3, 5).astype(F32)
    rtensor.sum(y * T(gy)).backward()
    out.update({"gap.x": x, "gap.gy": gy, "gap.y": y.numpy(), "gap.dx": xt.grad.numpy()})
    np.savez_compressed(os.path.join(GOLD, "ops.npz"), **out)


def golden_optim():
    out = {}
    rng = np.random.RandomState(4)
    shapes = [(7, 5), (1, 5), (3, 2, 3, 3)]
    for name, make in (("adam", lambda ps: Adam(ps, lr=5e-3, weight_decay=5e-4)),
                       ("adam_nowd", lambda ps: Adam(ps, lr=1e-3)),
                       ("sgd", lambda ps: SGD(ps, lr=0.05)),
                       ("sgd_mom", lambda ps: SGD(ps, lr=0.05, momentum=0.9, weight_decay=1e-3, nesterov=True))):
        fresh()
        p0 = [rng.randn(*s).astype(F32) for s in shapes]
        grads = [[rng.randn(*s).astype(F32) for s in shapes] for _ in range(3)]
        params = [T(p, True) for p in p0]
        opt = make(params)
        state = [(p.copy(), np.zeros_like(p), np.zeros_like(p)) for p in p0]
        for step, gs in enumerate(grads):
            for p, g in zip(params, gs):
                p.grad = rbt.BackendTensor(g, device=CPU)
            opt.step()
            for i, g in enumerate(gs):
                p, v, s = state[i]
                if name.startswith("adam"):
                    state[i] = ops.adam_step(p, g, v, s, opt.lr, opt.beta1, opt.beta2, opt.eps, opt.weight_decay, step + 1)
                else:
                    np_, nv = ops.sgd_step(p, g, v, opt.lr, opt.momentum, opt.weight_decay, opt.nesterov)
                    state[i] = (np_, nv if nv is not None else v, s)
        for i, p in enumerate(params):
            out["%s.p0.%d" % (name, i)] = p0[i]
            out["%s.p3.%d" % (name, i)] = p.numpy()
            close("%s param %d" % (name, i), state[i][0], p.numpy(), 1e-6)
            for st in range(3):
                out["%s.g%d.%d" % (name, st, i)] = grads[st][i]
    np.savez_compressed(os.path.join(GOLD, "optim.npz"), **out)


def train_steps(model, opt_factory, x, targets, steps, seed):
    """`steps` iterations of forward / loss / backward / step exactly as the scripts loop
    (test/ResNet_CIFAR10_cuda.py:185-201). Returns losses and logits of every step."""
    crit = rnn.CrossEntropyLoss()
    opt = opt_factory(model.parameters())
    losses, logits = [], []
    np.random.seed(seed)  # dropout masks
    model.train()
    for it in range(steps):
        xt, tt = T(x[it]), T(targets[it])
        outp = model(xt)
        loss = crit(outp, tt)
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(loss.data.numpy().item())
        logits.append(outp.data.numpy().copy())
        rtensor.Graph.free_graph()
    return np.array(losses, F32), np.stack(logits)


def golden_training():
    df = workloads.namespace(DeepFlows)
    rng = np.random.RandomState(5)

    def run(name, builder, opt_factory, in_shape, batch, steps=2, smooth=0.0):
        fresh()
        np.random.seed(11)
        model = builder()
        params0 = {k: p.numpy().copy() for k, p in workloads.all_parameters(model)}
        x = np.clip(rng.randn(steps, batch, *in_shape), -1, 1).astype(F32)
        tg = np.eye(10, dtype=F32)[rng.randint(0, 10, (steps, batch))]
        if smooth:
            tg = (tg * (1 - smooth) + smooth / 10).astype(F32)
        losses, logits = train_steps(model, opt_factory, x, tg, steps, seed=23)
        out = {"x": x, "target": tg, "losses": losses, "logits": logits}
        # Conditioning probe: replay the same steps with the inputs perturbed at float32 rounding level
        # (relative 1e-7). A parameter whose update moves by more than 2e-5 under that perturbation is
        # not reproducible by ANY implementation that rounds differently (typically: an Adam-normalised
        # step on a gradient that is analytically ~0, e.g. a bias feeding BatchNorm); the parity tests
        # only bound such parameters by the optimizer's step size.
        fresh()
        np.random.seed(11)
        twin = builder()
        xp = (x.astype(np.float64) * (1.0 + 1e-7 * np.sign(rng.randn(*x.shape)))).astype(F32)
        train_steps(twin, opt_factory, xp, tg, steps, seed=23)
        ill = []
        for (k, p), (_, q) in zip(workloads.all_parameters(model), workloads.all_parameters(twin)):
            a, b = p.numpy().astype(np.float64), q.numpy().astype(np.float64)
            if np.abs(a - b).max() / max(np.abs(a).max(), 1e-30) > 2e-5:
                ill.append(k)
        out["ill_conditioned"] = np.array(ill, dtype="U64")
        print("  ill-conditioned under 1e-7 input perturbation:", ill)
        for k, v in params0.items():
            out["p0." + k] = v
        for k, p in workloads.all_parameters(model):
            out["p1." + k] = p.numpy()
        for mod_name, mod in model.named_modules() if hasattr(model, "named_modules") else []:
            if hasattr(mod, "running_mean") and getattr(mod, "running_mean", None) is not None and hasattr(mod, "num_features"):
                out["rm." + mod_name] = mod.running_mean.numpy()
                out["rv." + mod_name] = mod.running_var.numpy()
        np.savez_compressed(os.path.join(GOLD, "train_%s.npz" % name), **out)
        print("  train_%s: losses %s" % (name, losses))

    run("mlp", lambda: workloads.mlp_mnist(df, "cpu", sizes=(64, 32, 16, 10)), lambda ps: SGD(ps, lr=0.05), (64,), 16,
        steps=3)
    run("cnn_mnist", lambda: workloads.cnn_mnist(df, "cpu", widths=(4, 8), in_hw=12), lambda ps: Adam(ps, lr=1e-3),
        (1, 12, 12), 4)
    run("cnn_cifar10", lambda: workloads.cnn_cifar10(df, "cpu", widths=(4, 8, 8), in_hw=16),
        lambda ps: Adam(ps, lr=5e-3, weight_decay=5e-4), (3, 16, 16), 4)
    run("resnet_registered", lambda: workloads.resnet_cifar(df, "cpu", widths=(4, 8, 8, 16), layers=(1, 1, 1, 1)),
        lambda ps: Adam(ps, lr=1e-3, weight_decay=5e-4), (3, 32, 32), 4, smooth=0.05)
    run("resnet_script", lambda: workloads.resnet_cifar(df, "cpu", widths=(4, 8, 8, 16), layers=(1, 1, 1, 1),
                                                        registered=False),
        lambda ps: Adam(ps, lr=1e-3, weight_decay=5e-4), (3, 32, 32), 4, smooth=0.05)


if __name__ == "__main__":
    for fn in (golden_l0, golden_conv, golden_bn_pool_act_loss, golden_optim, golden_training):
        print(fn.__name__)
        fn()
    print("golden fixtures written to", GOLD)
    for f in sorted(os.listdir(GOLD)):
        print("  %-28s %7.1f KB" % (f, os.path.getsize(os.path.join(GOLD, f)) / 1024))
