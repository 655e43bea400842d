"""Checkpoint format checks shared by the CPU and GPU tiers (reference: utils/model_utils.py:19-181)."""
import os

import numpy as np

import parity
import workloads

F32 = np.float32
REF_PKL = os.path.join(os.environ.get("DEEPFLOWS_REFERENCE", "/root/reference"), "test", "checkpoints-cifar10cuda_70%",
                       "cifar10_cnn_cuda_checkpoint.pkl")


def resume_equals_uninterrupted(device_name, tmp_path):
    """Two real Adam steps, save, a third step; then a FRESH model + optimizer, load, the same third step: parameters,
    both Adam moments and the step counter must come out bit-identical - i.e. v, s, t, lr and weight decay all round-trip
    with real (non-zero) contents, and BatchNorm's running statistics too."""
    from DeepFlows import backend_api, nn, tensor
    from DeepFlows.tensor import Tensor
    from DeepFlows.utils.model_utils import save_checkpoint, load_checkpoint
    df = parity.df_namespace()
    dev = backend_api.Device(device_name)
    rng = np.random.RandomState(5)
    batches = [(np.clip(rng.randn(8, 3, 16, 16), -1, 1).astype(F32), np.eye(10, dtype=F32)[rng.randint(0, 10, 8)]) for _ in range(3)]
    path = str(tmp_path / "resume.pkl")

    def build():
        tensor.Graph.free_graph_all()
        np.random.seed(9)
        model = workloads.cnn_cifar10(df, device_name, widths=(4, 8, 8), in_hw=16, dropout=0.0)
        return model, df.optim.Adam(model.parameters(), lr=2e-3, weight_decay=5e-4), nn.CrossEntropyLoss()

    def step(model, opt, crit, xb, tb):
        loss = crit(model(Tensor(xb, device=dev)), Tensor(tb, device=dev))
        opt.zero_grad()
        loss.backward()
        opt.step()
        tensor.Graph.free_graph()

    model, opt, crit = build()
    for xb, tb in batches[:2]:
        step(model, opt, crit, xb, tb)
    opt.lr = 1.5e-3                                    # a scheduler moved it
    save_checkpoint(model, opt, epoch=3, loss=0.5, save_path=path)
    assert any(np.abs(v.numpy()).max() > 0 for v in opt.v) and any(np.abs(s.numpy()).max() > 0 for s in opt.s)
    step(model, opt, crit, *batches[2])
    want = ({k: p.data.numpy().copy() for k, p in model.named_parameters()}, [v.numpy().copy() for v in opt.v],
            [s.numpy().copy() for s in opt.s], opt.t, {k: b.numpy().copy() for k, b in model.named_buffers()})

    model2, opt2, crit2 = build()
    for p in model2.parameters():                      # make sure nothing survives from construction
        p.data = p.data * 0.0
    info = load_checkpoint(model2, opt2, save_path=path)
    assert info["epoch"] == 3 and info["loss"] == 0.5 and opt2.lr == 1.5e-3 and opt2.t == 3
    step(model2, opt2, crit2, *batches[2])
    for k, p in model2.named_parameters():
        assert np.array_equal(p.data.numpy(), want[0][k]), k
    for a, b in zip(opt2.v, want[1]):
        assert np.array_equal(a.numpy(), b)
    for a, b in zip(opt2.s, want[2]):
        assert np.array_equal(a.numpy(), b)
    assert opt2.t == want[3]
    for k, b in model2.named_buffers():
        assert np.array_equal(b.numpy(), want[4][k]), k


def loads_reference_written_checkpoint(device_name):
    """The one full checkpoint the reference ships (written by its CNN_CIFAR10_cuda script on its own GPU build): every
    parameter, Adam's v / s / t and the hyper-parameters must load into the same model built here, and the model must run."""
    import pickle
    from DeepFlows import backend_api, tensor
    from DeepFlows.autograd import no_grad
    from DeepFlows.tensor import Tensor
    from DeepFlows.utils.model_utils import load_checkpoint
    raw = pickle.load(open(REF_PKL, "rb"))
    df = parity.df_namespace()
    tensor.Graph.free_graph_all()
    np.random.seed(0)
    model = workloads.cnn_cifar10(df, device_name)
    opt = df.optim.Adam(model.parameters(), lr=1.0, weight_decay=0.0)
    info = load_checkpoint(model, opt, save_path=REF_PKL)
    assert info["epoch"] == raw["epoch"] and info["loss"] == raw["loss"]
    names = [k for k, _ in model.named_parameters()]
    assert names == list(raw["model_parameters"]), "parameter names / order differ from the reference's"
    for k, p in model.named_parameters():
        assert np.array_equal(p.data.numpy(), raw["model_parameters"][k].astype(F32)), k
    st = raw["optimizer_state"]
    assert opt.lr == st["lr"] and opt.weight_decay == st["weight_decay"] and opt.t == st["t"]
    for i, (v, s) in enumerate(zip(opt.v, opt.s)):
        assert np.array_equal(v.numpy(), np.asarray(st["v"][i], F32)) and np.array_equal(s.numpy(), np.asarray(st["s"][i], F32)), i
    model.eval()
    with no_grad():
        out = model(Tensor(np.random.RandomState(1).randn(4, 3, 32, 32).astype(F32), device=backend_api.Device(device_name))).numpy()
    assert out.shape == (4, 10) and np.isfinite(out).all()
