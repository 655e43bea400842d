"""tests/golden/train_resnet_fixmean.npz: the registered ResNet of make_golden.py trained for two steps by the
REFERENCE with its `BackendTensor.mean` repaired (divide by the axis length instead of the total element count,
DeepFlows/backend/backend_tensor.py:659-662) - what `DEEPFLOWS_FIX_MEAN=1` selects in deepflows_b200.

Why a second ResNet fixture: with the quirk the two-step global average pool returns true_mean / (N^2 C^2 W), the
logits are the classifier bias, and every block gradient is ~1e-10 of the classifier's - the unrepaired fixture
cannot see an error in a block's conv dgrad / wgrad / BatchNorm backward. With the repair the gradients of all
layers are O(1e-2..1) and no conv weight is ill-conditioned.

Test infrastructure (build container only: imports the reference from /root/reference).
    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_fixmean.py
"""
import os
import sys

os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import numpy as np  # noqa: E402

import make_golden as mg  # noqa: E402  (sets up the reference import and its CUDA-semantics numpy device)

rbt = mg.rbt


def _fixed_mean(self, axis=None, keepdims=False):
    """The repair: sum / (length of the reduced axis); axis=None keeps sum / size (already correct)."""
    n = rbt.prod(self.shape) if axis is None else self.shape[axis if not isinstance(axis, (tuple, list)) else axis[0]]
    return self.sum(axis, keepdims=keepdims) / n


def main():
    import DeepFlows
    import workloads
    from DeepFlows.optim import Adam, SGD
    rbt.BackendTensor.mean = _fixed_mean
    df = workloads.namespace(DeepFlows)
    F32 = np.float32
    for name, optf, lr in (("resnet_fixmean", lambda ps: Adam(ps, lr=1e-3, weight_decay=5e-4), 1e-3),
                           ("resnet_fixmean_sgd", lambda ps: SGD(ps, lr=0.05, momentum=0.9, weight_decay=5e-4), 0.05)):
        rng = np.random.RandomState(7)
        steps, batch = 2, 4
        builder = lambda: workloads.resnet_cifar(df, "cpu", widths=(4, 8, 8, 16), layers=(1, 1, 1, 1))  # noqa: E731
        mg.fresh()
        np.random.seed(11)
        model = builder()
        params0 = {k: p.numpy().copy() for k, p in workloads.all_parameters(model)}
        x = np.clip(rng.randn(steps, batch, 3, 32, 32), -1, 1).astype(F32)
        tg = (np.eye(10, dtype=F32)[rng.randint(0, 10, (steps, batch))] * 0.95 + 0.005).astype(F32)
        # first-step gradients of every parameter (what the kernels are judged on), then the training steps
        crit = mg.rnn.CrossEntropyLoss()
        outp = model(mg.T(x[0]))
        loss = crit(outp, mg.T(tg[0]))
        loss.backward()
        grads = {k: p.grad.numpy().copy() for k, p in workloads.all_parameters(model) if p.grad is not None}
        for _, p in workloads.all_parameters(model):
            p.grad = None
        mg.fresh()
        np.random.seed(11)
        model = builder()  # running statistics restart
        losses, logits = mg.train_steps(model, optf, x, tg, steps, seed=23)
        out = {"x": x, "target": tg, "losses": losses, "logits": logits}
        mg.fresh()
        np.random.seed(11)
        twin = builder()
        xp = (x.astype(np.float64) * (1.0 + 1e-7 * np.sign(rng.randn(*x.shape)))).astype(F32)
        mg.train_steps(twin, optf, xp, tg, steps, seed=23)
        ill = []
        for (k, p), (_, q) in zip(workloads.all_parameters(model), workloads.all_parameters(twin)):
            a, b = p.numpy().astype(np.float64), q.numpy().astype(np.float64)
            if np.abs(a - b).max() / max(np.abs(a).max(), 1e-30) > 2e-5:
                ill.append(k)
        out["ill_conditioned"] = np.array(ill, dtype="U64")
        print(name, "losses", losses, "ill-conditioned:", ill)
        gmax = max(np.abs(v).max() for v in grads.values())
        print("  smallest max|grad| / largest:", min(np.abs(v).max() for v in grads.values()) / gmax)
        for k, v in params0.items():
            out["p0." + k] = v
        for k, v in grads.items():
            out["g0." + k] = v
        for k, p in workloads.all_parameters(model):
            out["p1." + k] = p.numpy()
        for mod_name, mod in model.named_modules():
            if getattr(mod, "running_mean", None) is not None and hasattr(mod, "num_features"):
                out["rm." + mod_name] = mod.running_mean.numpy()
                out["rv." + mod_name] = mod.running_var.numpy()
        np.savez_compressed(os.path.join(mg.GOLD, "train_%s.npz" % name), **out)


if __name__ == "__main__":
    main()
