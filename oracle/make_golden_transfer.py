"""Generates tests/golden/train_transfer.npz from the REFERENCE: the transfer-learning flow of SURVEY 8f rank 3
(reference: DeepFlows/nn/modules/module.py:471-542 `load_state_dict` / `load_weights`,
test/ResNet18_parameter_freezing_test.py:100-216 `freeze_model_layers` + `Adam(filter(lambda p: p.requires_grad, ...))`).

Runs only in the build container (imports the reference package through oracle/make_golden.py); usage:

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_transfer.py

Flow recorded: a "pretrained" ResNet gives its weights as {name: ndarray}; a second, differently initialised ResNet
takes all of them except the classifier through `load_weights` (strict=False), its stem and first stage are frozen
(`param.requires_grad = False`, the script's 'partial' strategy), Adam is built over the trainable parameters only,
and two training steps run. Stored: the weights dict, the fresh model's own initial values, the batches, losses,
logits, final parameters, BatchNorm running statistics and the names of the frozen parameters.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_golden as mg  # noqa: E402  (imports the reference, installs the CUDA-semantics numpy device)

import numpy as np  # noqa: E402
import workloads  # noqa: E402
from DeepFlows.optim import Adam  # noqa: E402  (the reference's)

F32 = np.float32
FROZEN_PREFIXES = ("conv1.", "bn1.", "layer1_0.")


def build(df, seed):
    mg.fresh()
    np.random.seed(seed)
    return workloads.resnet_cifar(df, "cpu", widths=(4, 8, 8, 16), layers=(1, 1, 1, 1))


def prepare(df, weights):
    """The fresh model with the donor's weights loaded and its stem / first stage frozen."""
    model = build(df, 32)
    init = {k: p.numpy().copy() for k, p in model.named_parameters()}
    model.load_weights(weights)
    frozen = []
    for k, p in model.named_parameters():
        if k.startswith(FROZEN_PREFIXES):
            p.requires_grad = False
            frozen.append(k)
    assert frozen and len(frozen) < len(init)
    return model, init, frozen


def optimizer(ps):
    return Adam(filter(lambda p: p.requires_grad, ps), lr=1e-3, weight_decay=5e-4)


def main():
    import DeepFlows
    df = workloads.namespace(DeepFlows)
    rng = np.random.RandomState(9)
    donor = build(df, 31)
    weights = {k: p.numpy().copy() for k, p in donor.named_parameters() if not k.startswith("fc.")}
    model, init, frozen = prepare(df, weights)
    steps, batch = 2, 4
    x = np.clip(rng.randn(steps, batch, 3, 32, 32), -1, 1).astype(F32)
    tg = (np.eye(10, dtype=F32)[rng.randint(0, 10, (steps, batch))] * 0.95 + 0.005).astype(F32)
    p0 = {k: p.numpy().copy() for k, p in model.named_parameters()}
    for k in weights:
        assert np.array_equal(p0[k], weights[k]), k          # load_weights took every offered tensor
    assert all(np.array_equal(p0[k], init[k]) for k in init if k.startswith("fc."))
    losses, logits = mg.train_steps(model, optimizer, x, tg, steps, seed=23)
    out = {"x": x, "target": tg, "losses": losses, "logits": logits, "frozen": np.array(frozen, dtype="U64")}
    # conditioning probe, as in make_golden.golden_training: the same steps with the inputs perturbed at float32
    # rounding level; parameters whose update moves by more than 2e-5 under that perturbation (Adam-normalised steps
    # on gradients that are analytically ~0, e.g. a BatchNorm bias that the next BatchNorm removes) cannot be
    # reproduced by any implementation that rounds differently and are only bounded by the step size in the tests
    twin, _, _ = prepare(df, weights)
    xp = (x.astype(np.float64) * (1.0 + 1e-7 * np.sign(rng.randn(*x.shape)))).astype(F32)
    mg.train_steps(twin, optimizer, xp, tg, steps, seed=23)
    ill = []
    for (k, p), (_, q) in zip(model.named_parameters(), twin.named_parameters()):
        a, b = p.numpy().astype(np.float64), q.numpy().astype(np.float64)
        if np.abs(a - b).max() / max(np.abs(a).max(), 1e-30) > 2e-5:
            ill.append(k)
    out["ill_conditioned"] = np.array(ill, dtype="U64")
    print("ill-conditioned under 1e-7 input perturbation:", ill)
    for k, v in weights.items():
        out["w." + k] = v
    for k, v in init.items():
        out["init." + k] = v
    moved = 0
    for k, p in model.named_parameters():
        out["p1." + k] = p.numpy()
        if k in frozen:
            assert np.array_equal(out["p1." + k], p0[k]), "frozen parameter %s moved in the reference" % k
        else:
            moved += int(not np.array_equal(out["p1." + k], p0[k]))
    assert moved == len(init) - len(frozen), "a trainable parameter did not move"
    for mod_name, mod in model.named_modules():
        if hasattr(mod, "num_features") and getattr(mod, "running_mean", None) is not None:
            out["rm." + mod_name] = mod.running_mean.numpy()
            out["rv." + mod_name] = mod.running_var.numpy()
    # strict loading errors of the reference, recorded as text so the host package can be held to the same wording
    try:
        model.load_state_dict(weights, strict=True)
        raise AssertionError("strict load of a partial dict must fail")
    except RuntimeError as e:
        out["strict_error"] = np.array(str(e))
    path = os.path.join(mg.GOLD, "train_transfer.npz")
    np.savez_compressed(path, **out)
    print("losses", losses, "frozen", frozen)
    print("strict error:", str(out["strict_error"])[:200])
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
