"""Shared drivers for the parity tests: run a workload on DeepFlows' host package (device `cpu` = the
oracle's numpy device, or `cuda` = libdfb200) and compare with the fixtures taken from the reference."""
import numpy as np

from conftest import golden, rel_err
import workloads

F32 = np.float32

TRAIN_CASES = {
    # name: (builder kwargs, optimizer, steps)
    "mlp": dict(build=lambda df, d: workloads.mlp_mnist(df, d, sizes=(64, 32, 16, 10)),
                opt=lambda optim, ps: optim.SGD(ps, lr=0.05)),
    "cnn_mnist": dict(build=lambda df, d: workloads.cnn_mnist(df, d, widths=(4, 8), in_hw=12),
                      opt=lambda optim, ps: optim.Adam(ps, lr=1e-3)),
    "cnn_cifar10": dict(build=lambda df, d: workloads.cnn_cifar10(df, d, widths=(4, 8, 8), in_hw=16),
                        opt=lambda optim, ps: optim.Adam(ps, lr=5e-3, weight_decay=5e-4), lr=5e-3),
    "resnet_registered": dict(build=lambda df, d: workloads.resnet_cifar(df, d, widths=(4, 8, 8, 16), layers=(1, 1, 1, 1)),
                              opt=lambda optim, ps: optim.Adam(ps, lr=1e-3, weight_decay=5e-4), lr=1e-3),
    "resnet_script": dict(build=lambda df, d: workloads.resnet_cifar(df, d, widths=(4, 8, 8, 16), layers=(1, 1, 1, 1),
                                                                    registered=False),
                          opt=lambda optim, ps: optim.Adam(ps, lr=1e-3, weight_decay=5e-4), lr=1e-3),
    # the reference with its mean(axis) repaired (oracle/make_golden_fixmean.py; DEEPFLOWS_FIX_MEAN=1 here): every block's
    # gradient is O(1), so these are the fixtures that see conv dgrad / wgrad / BatchNorm backward inside a real graph
    "resnet_fixmean": dict(build=lambda df, d: workloads.resnet_cifar(df, d, widths=(4, 8, 8, 16), layers=(1, 1, 1, 1)),
                           opt=lambda optim, ps: optim.Adam(ps, lr=1e-3, weight_decay=5e-4), lr=1e-3, fix_mean=True),
    "resnet_fixmean_sgd": dict(build=lambda df, d: workloads.resnet_cifar(df, d, widths=(4, 8, 8, 16), layers=(1, 1, 1, 1)),
                               opt=lambda optim, ps: optim.SGD(ps, lr=0.05, momentum=0.9, weight_decay=5e-4), lr=0.05,
                               fix_mean=True),
}


def df_namespace():
    import DeepFlows
    return workloads.namespace(DeepFlows)


def run_training_case(name, device_name):
    """Rebuild the model with the fixture's initial weights, replay the fixture's batches, return
    {losses, logits, params, running stats} plus the fixture."""
    import DeepFlows
    from DeepFlows import backend_api, tensor, nn
    from DeepFlows.tensor import Tensor
    g = golden("train_" + name)
    backend_api.set_dgrad_mode("reference")  # the fixtures hold the reference's last-writer-wins dgrad (conftest restores the default)
    df = df_namespace()
    case = TRAIN_CASES[name]
    dev = backend_api.Device(device_name)
    backend_api.set_fix_mean(bool(case.get("fix_mean")))
    try:
        np.random.seed(11)
        model = case["build"](df, device_name)
        for k, p in workloads.all_parameters(model):
            p.data = backend_api.Btensor(g["p0." + k], device=dev)
        opt = case["opt"](df.optim, model.parameters())
        crit = nn.CrossEntropyLoss()
        losses, logits, grads0 = [], [], {}
        np.random.seed(23)
        model.train()
        for it in range(g["x"].shape[0]):
            x, t = Tensor(g["x"][it], device=dev), Tensor(g["target"][it], device=dev)
            out = model(x)
            loss = crit(out, t)
            opt.zero_grad()
            loss.backward()
            if it == 0:
                grads0 = {k: p.grad.numpy().copy() for k, p in workloads.all_parameters(model) if p.grad is not None}
            opt.step()
            losses.append(loss.data.numpy().item())
            logits.append(out.data.numpy().copy())
            tensor.Graph.free_graph()
    finally:
        backend_api.set_fix_mean(False)
    params = {k: p.data.numpy() for k, p in workloads.all_parameters(model)}
    stats = {}
    for mod_name, mod in model.named_modules():
        if hasattr(mod, "num_features") and mod.running_mean is not None:
            stats["rm." + mod_name] = mod.running_mean.numpy()
            stats["rv." + mod_name] = mod.running_var.numpy()
    return dict(losses=np.array(losses, F32), logits=np.stack(logits), params=params, stats=stats, grads0=grads0), g


def check_training_case(name, device_name, tol_param=1e-4, tol_out=1e-4, tol_grad=1e-4):
    """north_star: one training step's parameter update within 1e-4 (relative to max |param|).

    Parameters listed in the fixture's `ill_conditioned` array are those whose update the REFERENCE ITSELF
    cannot reproduce when its inputs are perturbed at float32 rounding level (oracle/make_golden.py: an
    Adam-normalised step on a gradient that is analytically ~0, e.g. a bias feeding BatchNorm). They are
    only required to stay within the optimizer's step bound."""
    from DeepFlows import backend_api
    backend_api.set_dgrad_mode("reference")  # the fixtures hold the reference's last-writer-wins dgrad
    res, g = run_training_case(name, device_name)
    assert rel_err(res["losses"], g["losses"]) < tol_out, (res["losses"], g["losses"])
    assert rel_err(res["logits"], g["logits"]) < 5 * tol_out
    # first-step gradients (fixtures that hold them): each within tol of the reference, relative to its own magnitude
    # (floored at 1e-3 of the largest gradient of the net: below that a gradient is rounding noise of the others)
    g0 = {k[3:]: v for k, v in g.items() if k.startswith("g0.")}
    if g0:
        gmax = max(np.abs(v).max() for v in g0.values())
        ill0 = set(g.get("ill_conditioned", np.array([], dtype="U1")).tolist())
        for k, v in g0.items():
            # an analytically-zero gradient (a BatchNorm bias that the next BatchNorm removes) is rounding noise in the
            # reference itself: those are judged against the largest gradient of the net
            floor = gmax if k in ill0 else 1e-3 * gmax
            err = np.abs(res["grads0"][k].astype(np.float64) - v).max() / max(np.abs(v).max(), floor)
            assert err < tol_grad, "gradient of %s differs from the reference by %.3g" % (k, err)
    worst = 0.0
    case = TRAIN_CASES[name]
    for k, v in res["params"].items():
        if k in set(g.get("ill_conditioned", np.array([], dtype="U1")).tolist()):
            steps = g["x"].shape[0]
            assert np.abs(v - g["p0." + k]).max() <= 1.01 * case["lr"] * steps * (1 + np.abs(g["p0." + k]).max()) + 1e-6, k
            continue
        e = rel_err(v, g["p1." + k])
        worst = max(worst, e)
        assert e < tol_param, "parameter %s differs from the reference by %.3g" % (k, e)
        # and the step actually moved registered parameters
    # running statistics after the last step; when some parameters took ill-conditioned (noise-signed)
    # first steps, the second step's activations legitimately differ at the lr level
    tol_stats = tol_param if len(g.get("ill_conditioned", [])) == 0 else 2e-2
    for k, v in res["stats"].items():
        if k in g:
            assert rel_err(v, g[k]) < tol_stats, k
    return worst


# ---- transfer learning: fixture from oracle/make_golden_transfer.py -------------------------
FROZEN_PREFIXES = ("conv1.", "bn1.", "layer1_0.")


def run_transfer(device_name):
    from DeepFlows import backend_api, nn, tensor
    from DeepFlows.tensor import Tensor
    g = golden("train_transfer")
    backend_api.set_dgrad_mode("reference")  # fixture from the reference (last-writer-wins dgrad)
    df = df_namespace()
    dev = backend_api.Device(device_name)
    tensor.Graph.free_graph_all()
    np.random.seed(32)
    model = workloads.resnet_cifar(df, device_name, widths=(4, 8, 8, 16), layers=(1, 1, 1, 1))
    for k, p in model.named_parameters():                      # the fixture's own initial values (host RNG streams differ)
        p.data = backend_api.Btensor(g["init." + k], device=dev)
    weights = {k[2:]: v for k, v in g.items() if k.startswith("w.")}
    model.load_weights(weights)
    frozen = []
    for k, p in model.named_parameters():
        if k.startswith(FROZEN_PREFIXES):
            p.requires_grad = False
            frozen.append(k)
    p0 = {k: p.data.numpy().copy() for k, p in model.named_parameters()}
    opt = df.optim.Adam(filter(lambda p: p.requires_grad, model.parameters()), lr=1e-3, weight_decay=5e-4)
    crit = nn.CrossEntropyLoss()
    losses, logits = [], []
    np.random.seed(23)
    model.train()
    for it in range(g["x"].shape[0]):
        x, t = Tensor(g["x"][it], device=dev), Tensor(g["target"][it], device=dev)
        out = model(x)
        loss = crit(out, t)
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(loss.data.numpy().item())
        logits.append(out.data.numpy().copy())
        tensor.Graph.free_graph()
    return g, model, weights, frozen, p0, np.array(losses, np.float32), np.stack(logits), opt


def check_transfer(device_name, tol=1e-4):
    g, model, weights, frozen, p0, losses, logits, opt = run_transfer(device_name)
    assert frozen == g["frozen"].tolist()
    ill = set(g["ill_conditioned"].tolist())
    for k in weights:
        assert np.array_equal(p0[k], weights[k]), k             # load_weights took every offered tensor ...
    assert np.array_equal(p0["fc.weight"], g["init.fc.weight"])  # ... and left the rest alone
    assert len(opt.params) == len(p0) - len(frozen)
    assert rel_err(losses, g["losses"]) < tol and rel_err(logits, g["logits"]) < tol
    for k, p in model.named_parameters():
        got = p.data.numpy()
        if k in frozen:
            assert np.array_equal(got, p0[k]), "frozen parameter %s moved" % k
            assert p.grad is None, "frozen parameter %s received a gradient" % k
        elif k in ill:  # Adam-normalised step on a gradient that is ~0 analytically: bounded by the step size only
            assert np.abs(got - p0[k]).max() <= 1.01 * 1e-3 * g["x"].shape[0] * (1 + np.abs(p0[k]).max()) + 1e-6, k
        else:
            assert not np.array_equal(got, p0[k]), "trainable parameter %s did not move" % k
            assert rel_err(got, g["p1." + k]) < tol, k
    tol_stats = tol if not ill else 2e-2  # second-step activations see the ill-conditioned first steps at the lr level
    for mod_name, mod in model.named_modules():
        if hasattr(mod, "num_features") and mod.running_mean is not None:
            assert rel_err(mod.running_mean.numpy(), g["rm." + mod_name]) < tol_stats, mod_name
            assert rel_err(mod.running_var.numpy(), g["rv." + mod_name]) < tol_stats, mod_name
    return model, weights, g
