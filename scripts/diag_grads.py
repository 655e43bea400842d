"""Diagnostic (GPU box): per-parameter gradient error, cuda vs oracle numpy device, for the smoke model."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import deepflows_b200
from oracle import numpy_device
from DeepFlows import backend_api, nn, tensor
from DeepFlows.tensor import Tensor
import DeepFlows, workloads
backend_api.register_numpy_device(numpy_device)
backend_api.set_precision("fp32"); backend_api.set_dgrad_mode("exact")
df = workloads.namespace(DeepFlows)
rng = np.random.RandomState(0)
for label, build, bs in (("resnet-small", lambda d: workloads.resnet_cifar(df, d, widths=(8, 16, 16, 32), layers=(1, 1, 1, 1)), 8),
                         ("cnn_cifar", lambda d: workloads.cnn_cifar10(df, d, widths=(8, 16, 32), dropout=0.0), 8),
                         ("resnet18", lambda d: workloads.resnet_cifar(df, d), 4)):
    x = np.clip(rng.randn(bs, 3, 32, 32), -1, 1).astype(np.float32)
    t = (np.eye(10, dtype=np.float32)[rng.randint(0, 10, bs)] * 0.95 + 0.005).astype(np.float32)
    res = {}; init = None
    for name in ("cpu", "cuda"):
        tensor.Graph.free_graph_all(); np.random.seed(0)
        model = build(name); d = backend_api.Device(name)
        params = workloads.all_parameters(model)
        if init is None: init = [p.data.numpy().copy() for _, p in params]
        else:
            for (_, p), v in zip(params, init): p.data = backend_api.Btensor(v, device=d)
        logits = model(Tensor(x, device=d)); loss = nn.CrossEntropyLoss()(logits, Tensor(t, device=d))
        loss.backward()
        res[name] = (loss.data.numpy().item(), logits.data.numpy().copy(), [(k, p.grad.numpy().copy()) for k, p in params if p.grad is not None])
    print(label, "loss", res["cpu"][0], res["cuda"][0], "logit err", np.abs(res["cpu"][1]-res["cuda"][1]).max())
    gmax = max(np.abs(g).max() for _, g in res["cpu"][2])
    for (k, a), (_, b) in zip(res["cpu"][2], res["cuda"][2]):
        print("  %-28s max|g| %.3e  abs err %.3e  rel %.3e  (gmax %.3e)" % (k, np.abs(a).max(), np.abs(a-b).max(), np.abs(a-b).max()/max(np.abs(a).max(),1e-30), gmax))
