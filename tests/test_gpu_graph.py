"""CUDA-graph replay of a training step (DeepFlows.cuda_graph.CapturedStep) against the same steps run
eagerly: same kernels, same order, same buffers => bit-identical parameters, losses and optimizer state."""
import numpy as np
import pytest

import parity
import workloads

pytestmark = pytest.mark.gpu
F32 = np.float32


def _setup(df, opt_name, seed=5):
    from DeepFlows import backend_api, nn, tensor
    from DeepFlows.tensor import Tensor
    tensor.Graph.free_graph_all()
    np.random.seed(seed)
    model = workloads.resnet_cifar(df, "cuda", widths=(8, 16, 16, 32), layers=(1, 1, 1, 1), registered=True)
    if opt_name == "adam":
        opt = df.optim.Adam(model.parameters(), lr=1e-2, weight_decay=5e-4)
    else:
        opt = df.optim.SGD(model.parameters(), lr=0.05, momentum=0.9, weight_decay=1e-4, nesterov=True)
    return model, opt, nn.CrossEntropyLoss()


def _batches(n, batch=8):
    rng = np.random.RandomState(11)
    out = []
    for _ in range(n):
        x = np.clip(rng.randn(batch, 3, 32, 32), -1, 1).astype(F32)
        t = (np.eye(10, dtype=F32)[rng.randint(0, 10, batch)] * 0.95 + 0.005).astype(F32)
        out.append((x, t))
    return out


@pytest.mark.parametrize("opt_name", ["adam", "sgd"])
@pytest.mark.parametrize("precision", ["fp32", "tf32"])
def test_graph_replay_is_bit_identical_to_eager(cuda_device, opt_name, precision):
    from DeepFlows import backend_api, tensor
    from DeepFlows.tensor import Tensor
    from DeepFlows.cuda_graph import CapturedStep
    dev = cuda_device
    backend_api.set_precision(precision)
    backend_api.set_dgrad_mode("exact")
    try:
        df = parity.df_namespace()
        batches = _batches(6)
        lrs = [None, None, None, 0.5, None, 0.25]  # scheduler-style lr changes between steps (factor of the base lr)

        def run(graph):
            model, opt, crit = _setup(df, opt_name)
            base_lr = opt.lr
            x = Tensor(backend_api.Btensor(batches[0][0], device=dev))
            t = Tensor(backend_api.Btensor(batches[0][1], device=dev))

            def step_fn():
                loss = crit(model(x), t)
                opt.zero_grad()
                loss.backward()
                opt.step()
                return loss

            step = CapturedStep(step_fn, device=dev, warmup=1) if graph else step_fn
            losses = []
            for (xb, tb), f in zip(batches, lrs):
                if f is not None:
                    opt.lr = base_lr * f
                dev.from_numpy(xb, x.data._handle)
                dev.from_numpy(tb, t.data._handle)
                loss = step()
                losses.append(loss.data.numpy().copy())
            if graph:
                assert step.captured and step.calls == len(batches)
            params = [p.data.numpy().copy() for _, p in workloads.all_parameters(model)]
            state = [v.numpy().copy() for v in opt.v]
            if graph:
                step.destroy()
            return losses, params, state, opt

        l0 = dev.launch_count()
        e_losses, e_params, e_state, e_opt = run(False)
        eager_launches = dev.launch_count() - l0
        l0 = dev.launch_count()
        g_losses, g_params, g_state, g_opt = run(True)
        graph_launches = dev.launch_count() - l0
        assert graph_launches <= 0.55 * eager_launches  # 2 of the 6 steps issued kernels from Python (+ a few per replay: input uploads)
        for a, b in zip(e_losses, g_losses):
            assert np.array_equal(a, b)
        for a, b in zip(e_params, g_params):
            assert np.array_equal(a, b)
        for a, b in zip(e_state, g_state):
            assert np.array_equal(a, b)
        if opt_name == "adam":
            assert e_opt.t == g_opt.t
    finally:
        backend_api.set_precision("fp32")
        backend_api.set_dgrad_mode("exact")


def test_replays_launched_ahead_keep_their_own_hyper_parameters(cuda_device):
    """Replays queued back to back (the host never waits for the device, as in bench.py's timed loop): every replay must
    see ITS step's Adam bias corrections and learning rate. The hyper-parameter block lives in pinned memory and is read
    when the replay starts - the runtime makes the host wait for the previous replay's echo before it writes the next
    values (runtime.cu: GraphPool::hyper_ack); without that, step t ran with the values of step t + k."""
    from DeepFlows import backend_api, tensor
    from DeepFlows.tensor import Tensor
    from DeepFlows.cuda_graph import CapturedStep
    dev = cuda_device
    backend_api.set_precision("tf32")
    try:
        df = parity.df_namespace()
        xb, tb = _batches(1, batch=64)[0]
        steps = 40

        def run(graph):
            model, opt, crit = _setup(df, "adam")
            base_lr = opt.lr
            x = Tensor(backend_api.Btensor(xb, device=dev))
            t = Tensor(backend_api.Btensor(tb, device=dev))

            def step_fn():
                loss = crit(model(x), t)
                opt.zero_grad()
                loss.backward()
                opt.step()
                tensor.Graph.free_graph()
                return loss

            step = CapturedStep(step_fn, device=dev, warmup=1) if graph else step_fn
            for i in range(steps):
                opt.lr = base_lr * (0.5 if i % 7 == 3 else 1.0)   # changes between replays as well
                step()                                             # no read-back, no synchronisation in between
            dev.synchronize()
            params = [p.data.numpy().copy() for _, p in workloads.all_parameters(model)]
            if graph:
                step.destroy()
            return params

        eager, replayed = run(False), run(True)
        for a, b in zip(eager, replayed):   # (a wrong bias correction moves a parameter by ~1e-2 of the step; same kernels otherwise)
            assert np.abs(a - b).max() <= 1e-5 * max(np.abs(a).max(), 1e-3)
    finally:
        backend_api.set_precision("fp32")


def test_capture_rejects_host_copies(cuda_device):
    from DeepFlows import backend_api
    from DeepFlows.tensor import Tensor
    from DeepFlows.cuda_graph import CapturedStep
    dev = cuda_device
    x = Tensor(backend_api.Btensor(np.ones((4, 4), F32), device=dev))

    def bad():
        return (x * 2.0).numpy()

    step = CapturedStep(bad, device=dev, warmup=0)
    with pytest.raises(RuntimeError):
        step()
    assert not dev.graph_capturing()
    # the runtime is usable afterwards
    assert np.array_equal((x * 2.0).numpy(), np.full((4, 4), 2.0, F32))


def test_graph_pool_keeps_captured_buffers_private(cuda_device):
    """Blocks used by a captured step must not be handed to later eager allocations."""
    from DeepFlows import backend_api
    from DeepFlows.tensor import Tensor
    from DeepFlows.cuda_graph import CapturedStep
    dev = cuda_device
    x = Tensor(backend_api.Btensor(np.arange(1024, dtype=F32), device=dev))
    holder = {}

    def fn():
        holder["y"] = (x * 3.0) + 1.0  # one temporary + one result
        return holder["y"]

    step = CapturedStep(fn, device=dev, warmup=1)
    step()
    y = step()
    want = np.arange(1024, dtype=F32) * 3 + 1
    assert np.array_equal(y.numpy(), want)
    # allocate and scribble over many same-sized eager buffers, then replay
    junk = [backend_api.full((1024,), -7.0, device=dev) for _ in range(16)]
    dev.from_numpy(np.arange(1024, dtype=F32)[::-1].copy(), x.data._handle)
    y2 = step()
    assert np.array_equal(y2.numpy(), np.arange(1024, dtype=F32)[::-1] * 3 + 1)
    assert all(np.array_equal(j.numpy(), np.full(1024, -7.0, F32)) for j in junk)
    step.destroy()


def test_device_prefetcher_delivers_every_batch(cuda_device):
    """utils.data.DevicePrefetcher: batches arrive on the device through pinned memory and the copy stream, in
    order, bit-exact, including a ragged last batch; `into=` refreshes static tensors (the CapturedStep pattern)."""
    from DeepFlows import backend_api
    from DeepFlows.tensor import Tensor
    from DeepFlows.utils.data import data_loader, DevicePrefetcher
    rng = np.random.RandomState(2)
    X = rng.randn(70, 3, 8, 8).astype(F32)
    Y = np.eye(10, dtype=F32)[rng.randint(0, 10, 70)]
    loader = data_loader(X, Y, batch_size=16)            # 16, 16, 16, 16, 6
    seen = 0
    for x, t in DevicePrefetcher(loader, device=cuda_device):
        n = x.shape[0]
        junk = backend_api.full((4096,), float(seen), device=cuda_device)  # keep the compute stream busy meanwhile
        assert np.array_equal(x.numpy(), X[seen:seen + n]) and np.array_equal(t.numpy(), Y[seen:seen + n])
        assert np.array_equal(junk.numpy(), np.full(4096, float(seen), F32))
        seen += n
    assert seen == 70
    loader = data_loader(X[:64], Y[:64], batch_size=16)
    xs = Tensor(backend_api.Btensor(np.zeros((16, 3, 8, 8), F32), device=cuda_device))
    ts = Tensor(backend_api.Btensor(np.zeros((16, 10), F32), device=cuda_device))
    for i, (x, t) in enumerate(DevicePrefetcher(loader, device=cuda_device, into=(xs, ts))):
        assert x is xs and t is ts
        assert np.array_equal(xs.numpy(), X[16 * i:16 * i + 16]) and np.array_equal(ts.numpy(), Y[16 * i:16 * i + 16])
    assert i == 3
