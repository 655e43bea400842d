"""Host-side logic (BackendTensor views, the autograd tape, nn / optim, fused-op wiring) on the oracle's
numpy device, against fixtures produced by the reference package (oracle/make_golden.py). No GPU."""
import numpy as np
import pytest

from conftest import golden, rel_err
import parity

F32 = np.float32


@pytest.mark.parametrize("name", list(parity.TRAIN_CASES))
def test_training_steps_match_reference(cpu_device, name):
    parity.check_training_case(name, "cpu")


def test_script_variant_registers_five_parameters(cpu_device):
    # blocks held in Python lists are invisible to parameters()
    import workloads
    df = parity.df_namespace()
    m = workloads.resnet_cifar(df, "cpu", widths=(4, 8, 8, 16), layers=(1, 1, 1, 1), registered=False)
    assert [n for n, _ in m.named_parameters()] == ["conv1.weight", "bn1.weight", "bn1.bias", "fc.weight", "fc.bias"]
    m2 = workloads.resnet_cifar(df, "cpu", widths=(4, 8, 8, 16), layers=(1, 1, 1, 1), registered=True)
    assert len(list(m2.parameters())) == len(workloads.all_parameters(m2)) > 5


def test_view_algebra_matches_reference(cpu_device):
    from DeepFlows import backend_api
    g = golden("l0")
    x = backend_api.Btensor(g["x"], device=cpu_device)
    assert np.array_equal(x.permute((2, 0, 3, 1)).compact().numpy(), g["permute_compact"])
    assert np.array_equal(x[1:3, 0:4:2, 2, 1:6:2].compact().numpy(), g["slice_compact"])
    assert np.array_equal(backend_api.Btensor(g["x"][:1, :, :1, :], device=cpu_device).broadcast_to((3, 4, 5, 6)).compact().numpy(),
                          g["bcast"])
    assert np.array_equal(x.pad(((0, 0), (0, 0), (2, 2), (1, 1))).numpy(), g["pad"])
    z = cpu_device.full((4, 6), 0.0)
    z[1:3, 0:6:2] = backend_api.Btensor(np.arange(6, dtype=F32).reshape(2, 3), device=cpu_device)
    z[3, 1:5] = 7.0
    assert np.array_equal(z.numpy(), g["setitem"])
    assert rel_err(x.sum(axis=1).numpy(), g["sum_axis1"]) < 1e-6
    assert np.array_equal(x.max(axis=2, keepdims=True).numpy(), g["max_axis2"])
    assert rel_err(x.mean(axis=2).numpy(), g["mean_axis2_quirk"]) < 1e-6
    m = backend_api.Btensor(g["m1"], device=cpu_device) @ backend_api.Btensor(g["m2"], device=cpu_device)
    assert rel_err(m.numpy(), g["matmul"]) < 1e-6
    with pytest.raises(ValueError):
        x.permute((1, 0, 2, 3)).reshape((60, 6))  # not compact


def test_dense_fast_path_keeps_channels_last(cpu_device):
    from DeepFlows import backend_api
    x = np.random.RandomState(0).randn(2, 3, 4, 5).astype(F32)
    t = backend_api.Btensor(x, device=cpu_device).channels_last()
    assert t.is_channels_last() and not t.is_compact() and t.is_dense()
    y = t * 2.0 + t
    assert y.strides == t.strides
    assert np.array_equal(y.numpy(), x * 2 + x)
    b = backend_api.Btensor(np.arange(3, dtype=F32).reshape(1, 3, 1, 1), device=cpu_device)
    assert np.array_equal((t + b).numpy(), x + np.arange(3, dtype=F32).reshape(1, 3, 1, 1))


def test_autograd_ops_against_numpy(cpu_device):
    from DeepFlows import tensor
    from DeepFlows.tensor import Tensor
    rng = np.random.RandomState(1)
    a, b = rng.rand(4, 5).astype(F32) + 0.5, rng.rand(4, 5).astype(F32) + 0.5
    ta, tb = Tensor(a, device=cpu_device, requires_grad=True), Tensor(b, device=cpu_device, requires_grad=True)
    out = tensor.sum(tensor.log(ta * tb + 1.0) / tb - tensor.exp(ta) * 0.1 + ta ** 2.0 + tensor.maximum(ta, 1.0))
    out.backward()
    ga = b / (a * b + 1) / b - np.exp(a) * 0.1 + 2 * a + (np.maximum(a, 1) == a)
    gb = (a / (a * b + 1)) / b - np.log(a * b + 1) / b ** 2
    assert rel_err(ta.grad.numpy(), ga) < 1e-5 and rel_err(tb.grad.numpy(), gb) < 1e-5
    # broadcast gradient is reduced on the device
    tensor.Graph.free_graph_all()
    w = Tensor(rng.randn(1, 5).astype(F32), device=cpu_device, requires_grad=True)
    x = Tensor(a, device=cpu_device)
    tensor.sum((x + w) * x).backward()
    assert rel_err(w.grad.numpy(), a.sum(axis=0, keepdims=True)) < 1e-6


def test_gap_quirk_q3(cpu_device):
    from DeepFlows import tensor
    from DeepFlows.tensor import Tensor
    g = golden("ops")
    x = Tensor(g["gap.x"], device=cpu_device, requires_grad=True)
    y = tensor.mean(tensor.mean(x, axis=2), axis=2)
    tensor.sum(y * Tensor(g["gap.gy"], device=cpu_device)).backward()
    assert rel_err(y.numpy(), g["gap.y"]) < 1e-6 and rel_err(x.grad.numpy(), g["gap.dx"]) < 1e-6


def test_module_registry_and_state(cpu_device, tmp_path):
    from DeepFlows import nn, backend_api
    from DeepFlows.optim import Adam
    from DeepFlows.utils.model_utils import save_checkpoint, load_checkpoint
    import workloads
    df = parity.df_namespace()
    m = workloads.cnn_cifar10(df, "cpu", widths=(4, 8, 8), in_hw=16)
    names = [n for n, _ in m.named_parameters()]
    assert names[:4] == ["conv1.weight", "conv1.bias", "bn1.weight", "bn1.bias"] and names[-2:] == ["fc.weight", "fc.bias"]
    assert [n for n, _ in m.named_buffers()][:2] == ["bn1.running_mean", "bn1.running_var"]
    opt = Adam(m.parameters(), lr=1e-3)
    opt.t = 7
    path = str(tmp_path / "ck.pkl")
    save_checkpoint(m, opt, epoch=3, loss=0.5, save_path=path)
    before = {k: v.copy() for k, v in m.state_dict().items()}
    for p in m.parameters():
        p.data = p.data * 0.0
    opt.t = 1
    info = load_checkpoint(m, opt, save_path=path)
    assert info == {"epoch": 3, "loss": 0.5} and opt.t == 7
    for k, v in m.state_dict().items():
        assert np.array_equal(v, before[k])
    m.eval()
    from DeepFlows import autograd
    assert not autograd.is_grad_enable()  # quirk Q9
    m.train()
    assert autograd.is_grad_enable()


def test_dropout_eval_quirk_q8(cpu_device):
    from DeepFlows import nn
    from DeepFlows.tensor import Tensor
    d = nn.Dropout(0.25)
    x = Tensor(np.ones((2, 4), F32), device=cpu_device)
    d.eval()
    assert np.allclose(d(x).numpy(), 0.75)
    d.train()
    np.random.seed(0)
    y = d(x).numpy()
    assert set(np.unique(y)).issubset({0.0, np.float32(1 / 0.75)})


def test_schedulers():
    from DeepFlows.optim.scheduler import StepLR, CosineAnnealingLR, WarmupCosineLR

    class O:
        lr = 1.0
    o = O()
    s = WarmupCosineLR(o, warmup_epochs=2, T_max=4, eta_min=0.1)
    lrs = []
    for _ in range(5):
        s.step()
        lrs.append(o.lr)
    assert lrs[0] == 0.0 and abs(lrs[1] - 0.5) < 1e-12 and abs(lrs[2] - 1.0) < 1e-12 and lrs[3] < 1.0
    o2 = O()
    c = CosineAnnealingLR(o2, T_max=10, eta_min=0.0)
    c.step()
    assert abs(o2.lr - 1.0) < 1e-12
    o3 = O()
    st = StepLR(o3, step_size=2, gamma=0.5)
    for _ in range(3):
        st.step()
    assert abs(o3.lr - 0.5) < 1e-12


def test_memory_layout_helpers_and_optimizers_follow_parameter_layout(cpu_device):
    """Conv weights live channels-last (K,R,R,C) behind their logical (K,C,R,R) shape on the B200 device. The host
    logic that makes this invisible - `with_layout_of`, `flat_storage`, the fused optimizers bringing gradient and
    state to the parameter's memory layout - is device-independent and checked here on the numpy device."""
    from DeepFlows import backend_api
    from DeepFlows.tensor import Tensor
    from DeepFlows.optim import Adam, SGD
    rng = np.random.RandomState(3)
    w = rng.randn(6, 4, 3, 3).astype(F32)
    g = rng.randn(6, 4, 3, 3).astype(F32)
    compact = backend_api.Btensor(w, device=cpu_device)
    cl = compact.channels_last()
    assert cl.is_channels_last() and cl.is_dense() and not cl.is_compact()
    assert np.array_equal(cl.numpy(), w)
    assert np.array_equal(cl.flat_storage().numpy(), w.transpose(0, 2, 3, 1).reshape(-1))      # memory order
    gt = backend_api.Btensor(g, device=cpu_device)
    g_cl = gt.with_layout_of(cl)
    assert g_cl.strides == cl.strides and np.array_equal(g_cl.numpy(), g)
    assert gt.with_layout_of(compact) is gt and g_cl.with_layout_of(cl) is g_cl              # no copy when layouts agree
    assert np.array_equal(g_cl.with_layout_of(compact).numpy(), g) and g_cl.with_layout_of(compact).is_compact()
    sliced = backend_api.Btensor(rng.randn(6, 8, 3, 3).astype(F32), device=cpu_device)[:, 0:8:2, :, :]  # neither compact nor dense
    assert np.array_equal(sliced.with_layout_of(cl).numpy(), sliced.numpy())

    for make in (lambda ps: Adam(ps, lr=1e-2, weight_decay=1e-3), lambda ps: SGD(ps, lr=0.1, momentum=0.9, nesterov=True)):
        results = []
        for layout in ("compact", "channels_last", "switch"):
            p = Tensor(backend_api.Btensor(w, device=cpu_device), requires_grad=True)
            if layout == "channels_last":
                p.data = p.data.channels_last()
            opt = make([p])
            for step in range(3):
                if layout == "switch" and step == 1:       # e.g. a checkpoint was loaded, or the first conv re-laid it out
                    p.data = p.data.channels_last()
                p.grad = backend_api.Btensor(g * (step + 1), device=cpu_device)  # gradient arrives compact
                opt.step()
            results.append(p.data.numpy().copy())
            assert all(s.strides == p.data.strides for s in opt.v)
        assert np.array_equal(results[0], results[1]) and np.array_equal(results[0], results[2])


def test_captured_step_protocol_with_a_mock_device():
    """DeepFlows.cuda_graph.CapturedStep: call 1..warmup run eagerly, the next call captures AND launches (a capture
    executes nothing), later calls refresh the hyper-parameters of every optimizer that stepped during the capture
    (in capture order) and replay; a failing capture is torn down and re-raised."""
    from DeepFlows.cuda_graph import CapturedStep, note_optimizer_step, capturing

    class MockDevice:
        def __init__(self):
            self.log = []

        def graph_begin_capture(self):
            self.log.append("begin")

        def graph_end_capture(self):
            self.log.append("end")
            return 42

        def graph_launch(self, g):
            self.log.append(("launch", g))

        def graph_destroy(self, g):
            self.log.append(("destroy", g))

        def graph_node_counts(self, g):
            return (7, 9)

    class MockOpt:
        def __init__(self, name):
            self.name, self.t = name, 1

        def step(self):
            note_optimizer_step(self)
            self.t += 1

        def _graph_refresh(self, dev, g, index):
            dev.log.append(("refresh", self.name, index, self.t))
            self.t += 1

    dev, a, b = MockDevice(), MockOpt("a"), MockOpt("b")
    calls = []

    def fn():
        calls.append(capturing())
        a.step()
        b.step()
        return "loss"

    step = CapturedStep(fn, device=dev, warmup=2)
    assert step() == "loss" and step() == "loss" and not step.captured and dev.log == []
    assert step() == "loss" and step.captured
    assert dev.log == ["begin", "end", ("launch", 42)] and calls == [False, False, True]
    assert step.node_counts() == (7, 9)
    assert step() == "loss" and step() == "loss"
    assert dev.log[3:] == [("refresh", "a", 0, 4), ("refresh", "b", 1, 4), ("launch", 42),
                           ("refresh", "a", 0, 5), ("refresh", "b", 1, 5), ("launch", 42)]
    assert len(calls) == 3 and a.t == 6                        # replays do not run the Python step
    step.destroy()
    assert dev.log[-1] == ("destroy", 42) and not step.captured

    dev2 = MockDevice()

    def bad():
        raise ValueError("boom")

    failing = CapturedStep(bad, device=dev2, warmup=0)
    with pytest.raises(ValueError):
        failing()
    assert dev2.log == ["begin", "end", ("destroy", 42)] and not capturing() and not failing.captured


def test_lr_schedulers_closed_form():
    """optim/scheduler.py against the formulas of the reference (scheduler.py:14-59), values worked out by hand."""
    import math
    from DeepFlows.optim.scheduler import StepLR, CosineAnnealingLR, WarmupCosineLR

    class Opt:
        def __init__(self, lr):
            self.lr = lr

    o = Opt(0.1)
    s = StepLR(o, step_size=2, gamma=0.5)
    seen = []
    for _ in range(5):
        s.step()
        seen.append(o.lr)
    assert seen == [0.1, 0.1, 0.05, 0.05, 0.025]          # epoch 0 never decays
    o = Opt(1.0)
    s = CosineAnnealingLR(o, T_max=4, eta_min=0.0)
    seen = []
    for _ in range(6):
        s.step()
        seen.append(o.lr)
    want = [(1 + math.cos(math.pi * (e % 4) / 4)) / 2 for e in range(6)]
    assert seen == want and seen[4] == 1.0                   # restarts at T_max
    o = Opt(1e-3)
    s = WarmupCosineLR(o, warmup_epochs=5, T_max=20, eta_min=1e-5)   # the ResNet script's schedule
    seen = []
    for _ in range(8):
        s.step()
        seen.append(o.lr)
    assert seen[0] == 0.0 and seen[5] == 1e-3 and abs(seen[2] - 4e-4) < 1e-18
    assert seen[7] == 1e-5 + (1e-3 - 1e-5) * (1 + math.cos(math.pi * 2 / 20)) / 2

    class NoLr:
        pass
    n = NoLr()
    CosineAnnealingLR(n, 3).step()
    StepLR(n, 1).step()
    assert not hasattr(n, "lr")
