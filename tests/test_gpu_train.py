"""End-to-end parity on a B200 through the unchanged DeepFlows API (`device='cuda'`): training steps of
the script models against fixtures from the reference, and size-independent properties at the
BASELINE.json sizes (ResNet-18 CIFAR shape, batch 256)."""
import numpy as np
import pytest

from conftest import rel_err
import parity
import workloads

pytestmark = pytest.mark.gpu
F32 = np.float32


@pytest.mark.parametrize("name", list(parity.TRAIN_CASES))
def test_training_steps_match_reference(cuda_device, name):
    from DeepFlows import backend_api
    backend_api.set_precision("fp32")
    parity.check_training_case(name, "cuda")  # parameter update within 1e-4 (north_star)


@pytest.mark.parametrize("precision", ["tf32", "bf16"])
def test_training_steps_reduced_precision(cuda_device, precision):
    from DeepFlows import backend_api
    backend_api.set_precision(precision)
    try:
        res, g = parity.run_training_case("cnn_mnist", "cuda")
        assert rel_err(res["losses"], g["losses"]) < 2e-2
        assert rel_err(res["logits"], g["logits"]) < 2e-2
    finally:
        backend_api.set_precision("fp32")


def test_cuda_matches_numpy_device_on_resnet18_shapes(cuda_device, cpu_device):
    """Full-width ResNet-18 (32-64-128-256), 32x32 input, small batch: cuda vs the oracle device, same
    host code, exact dgrad."""
    from DeepFlows import backend_api, nn, tensor
    from DeepFlows.tensor import Tensor
    from DeepFlows.optim import SGD
    backend_api.set_precision("fp32")
    backend_api.set_dgrad_mode("exact")
    try:
        df = parity.df_namespace()
        rng = np.random.RandomState(0)
        x = np.clip(rng.randn(4, 3, 32, 32), -1, 1).astype(F32)
        t = (np.eye(10, dtype=F32)[rng.randint(0, 10, 4)] * 0.95 + 0.005).astype(F32)
        results = {}
        for dev_name in ("cpu", "cuda"):
            tensor.Graph.free_graph_all()
            np.random.seed(3)
            model = workloads.resnet_cifar(df, dev_name)
            dev = backend_api.Device(dev_name)
            if dev_name == "cuda":  # same initial weights on both devices
                for (k, p), (_, q) in zip(workloads.all_parameters(model), results["cpu_params0"]):
                    p.data = backend_api.Btensor(q, device=dev)
            else:
                results["cpu_params0"] = [(k, p.data.numpy().copy()) for k, p in workloads.all_parameters(model)]
            opt = SGD(model.parameters(), lr=0.1)
            out = model(Tensor(x, device=dev))
            loss = nn.CrossEntropyLoss()(out, Tensor(t, device=dev))
            opt.zero_grad()
            loss.backward()
            grads = {k: p.grad.numpy().copy() for k, p in workloads.all_parameters(model) if p.grad is not None}
            opt.step()
            results[dev_name] = (loss.data.numpy().item(), out.data.numpy().copy(), grads)
        assert abs(results["cpu"][0] - results["cuda"][0]) < 1e-5
        assert rel_err(results["cuda"][1], results["cpu"][1]) < 1e-4
        gmax = max(np.abs(v).max() for v in results["cpu"][2].values())
        for k, v in results["cpu"][2].items():
            # SGD-level comparison: gradient error relative to the largest gradient in the net
            assert np.abs(results["cuda"][2][k] - v).max() <= 1e-4 * max(np.abs(v).max(), 1e-3 * gmax), k
    finally:
        backend_api.set_dgrad_mode("exact")


def test_full_size_properties_resnet18_batch256(cuda_device):
    """BASELINE.json config 4 at full size. The oracle cannot run this in seconds, so check
    size-independent properties: conv adjointness <conv(x), gy> == <x, dgrad(gy)> == <w, wgrad(x, gy)>,
    linearity of fprop, BN output statistics, pool / ReLU idempotence, loss = ln(10) at init."""
    from DeepFlows import backend_api, nn, tensor
    from DeepFlows.nn import functional as F
    from DeepFlows.tensor import Tensor
    backend_api.set_precision("fp32")
    backend_api.set_dgrad_mode("exact")
    try:
        rng = np.random.RandomState(0)
        dev = cuda_device
        for (c, h, k, r, p, s) in [(32, 16, 32, 3, 1, 1), (32, 16, 64, 3, 1, 2), (64, 8, 128, 1, 0, 2), (3, 32, 32, 3, 1, 1)]:
            tensor.Graph.free_graph_all()
            x = rng.randn(256, c, h, h).astype(F32)
            w = (rng.randn(k, c, r, r) / np.sqrt(c * r * r)).astype(F32)
            xt, wt = Tensor(x, device=dev, requires_grad=True), Tensor(w, device=dev, requires_grad=True)
            y = F.conv2d(xt, wt, p, s)
            yv = y.numpy()
            gy = rng.randn(*yv.shape).astype(F32)
            tensor.sum(y * Tensor(gy, device=dev)).backward()
            lhs = float((yv.astype(np.float64) * gy).sum())
            assert abs(lhs - float((xt.grad.numpy().astype(np.float64) * x).sum())) < 1e-4 * abs(lhs) + 1e-2
            assert abs(lhs - float((wt.grad.numpy().astype(np.float64) * w).sum())) < 1e-4 * abs(lhs) + 1e-2
            y2 = F.conv2d(Tensor(2 * x, device=dev), Tensor(w, device=dev), p, s).numpy()
            assert rel_err(y2, 2 * yv) < 1e-6
        # BN output has zero mean / unit variance per channel; ReLU and MaxPool-of-constant are idempotent
        tensor.Graph.free_graph_all()
        x = (rng.randn(256, 32, 32, 32) * 3 + 5).astype(F32)
        bn = nn.BatchNorm2d(32, device="cuda")
        y = bn(Tensor(x, device=dev)).numpy()
        assert np.abs(y.mean(axis=(0, 2, 3))).max() < 1e-4 and np.abs(y.var(axis=(0, 2, 3)) - 1).max() < 1e-3
        r1 = F.relu(Tensor(y, device=dev))
        assert np.array_equal(F.relu(r1).numpy(), r1.numpy()) and np.array_equal(r1.numpy(), np.maximum(y, 0))
        pooled = F.max_pool2d(Tensor(y, device=dev), 2, 2).numpy()
        assert np.array_equal(pooled, y.reshape(256, 32, 16, 2, 16, 2).max(axis=(3, 5)))
        # one full training step of the script model at batch 256
        tensor.Graph.free_graph_all()
        df = parity.df_namespace()
        np.random.seed(0)
        model = workloads.resnet_cifar(df, "cuda")
        opt = df.optim.Adam(model.parameters(), lr=1e-3, weight_decay=5e-4)
        xb = np.clip(rng.randn(256, 3, 32, 32), -1, 1).astype(F32)
        tb = (np.eye(10, dtype=F32)[rng.randint(0, 10, 256)] * 0.95 + 0.005).astype(F32)
        before = {k: p.data.numpy().copy() for k, p in model.named_parameters()}
        loss = nn.CrossEntropyLoss()(model(Tensor(xb, device=dev)), Tensor(tb, device=dev))
        opt.zero_grad()
        loss.backward()
        opt.step()
        lv = loss.data.numpy().item()
        assert np.isfinite(lv) and abs(lv - np.log(10)) < 0.35  # Q3: logits ~ fc bias only at init
        moved = [np.abs(p.data.numpy() - before[k]).max() for k, p in model.named_parameters()]
        assert all(np.isfinite(mv) for mv in moved) and max(moved) > 0 and max(moved) <= 1.01e-3 * 1.5
    finally:
        backend_api.set_dgrad_mode("exact")


def test_checkpoint_roundtrip_on_cuda(cuda_device, tmp_path):
    from DeepFlows.utils.model_utils import save_checkpoint, load_checkpoint
    from DeepFlows.optim import Adam
    df = parity.df_namespace()
    np.random.seed(1)
    model = workloads.cnn_mnist(df, "cuda", widths=(4, 8), in_hw=12)
    opt = Adam(model.parameters(), lr=1e-3)
    path = str(tmp_path / "c.pkl")
    save_checkpoint(model, opt, epoch=1, loss=1.0, save_path=path)
    saved = {k: v.copy() for k, v in model.state_dict().items()}
    for p in model.parameters():
        p.data = p.data * 0.0
    load_checkpoint(model, opt, save_path=path)
    for k, v in model.state_dict().items():
        assert np.array_equal(v, saved[k])


def test_checkpoint_resume_equals_uninterrupted_on_cuda(cuda_device, tmp_path):
    """Adam's v / s / t and BatchNorm's running statistics after REAL steps round-trip through the reference's format."""
    import ckpt_checks
    from DeepFlows import backend_api
    backend_api.set_precision("fp32")
    ckpt_checks.resume_equals_uninterrupted("cuda", tmp_path)


def test_transfer_learning_matches_reference_on_cuda(cuda_device):
    """SURVEY 8f rank 3 on the device: `load_weights` of a partial dict, frozen stem / first stage (their conv
    weights still get the channels-last layout on first use, no wgrad is launched for them), Adam over the trainable
    subset; same fixture and bounds as tests/test_transfer_cpu.py."""
    from DeepFlows import backend_api
    backend_api.set_precision("fp32")
    parity.check_transfer("cuda")
