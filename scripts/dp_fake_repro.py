#!/usr/bin/env python
"""Single-GPU reproduction of the data-parallel host path: a fake transport with world = 2 whose all-reduce doubles the
bucket (as if the other rank held the same shard), so the data-parallel step must equal the plain single-process step."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import deepflows_b200  # noqa: E402,F401
import DeepFlows  # noqa: E402
import workloads  # noqa: E402
from DeepFlows import backend_api, dist, nn, tensor  # noqa: E402
from DeepFlows.tensor import Tensor  # noqa: E402

F32 = np.float32
dev = backend_api.cuda()
backend_api.set_precision(os.environ.get("PREC", "fp32"))
backend_api.set_dgrad_mode("exact")
df = workloads.namespace(DeepFlows)
B = 8
rng = np.random.RandomState(7)
X = np.clip(rng.randn(B, 3, 32, 32), -1, 1).astype(F32)
T = (np.eye(10, dtype=F32)[rng.randint(0, 10, B)] * 0.95 + 0.005).astype(F32)


class FakeTransport:
    rank, world = 0, 2

    def allreduce_sum(self, flat):
        dev.scalar_mul((flat._handle, flat._offset) if False else flat._handle, 2.0, flat._handle) if flat._offset == 0 and flat.size == flat._handle.size else flat.__setitem__(slice(None), flat * 2.0)

    def broadcast(self, flat, root=0):
        pass

    def wait(self):
        pass

    def close(self):
        pass


def build(seed):
    tensor.Graph.free_graph_all()
    np.random.seed(seed)
    return workloads.resnet_cifar(df, "cuda", widths=(8, 16, 16, 32), layers=(1, 1, 1, 1), registered=True)


def one_step(model, use_dp):
    opt = df.optim.SGD(model.parameters(), lr=0.1, momentum=0.9, weight_decay=5e-4)
    loss = nn.CrossEntropyLoss()(model(Tensor(X, device=dev)), Tensor(T, device=dev))
    opt.zero_grad()
    loss.backward()
    scale = dist.pre_step()
    grads = [p.grad.numpy().copy() * scale for p in model.parameters()]
    opt.step()
    tensor.Graph.free_graph()
    return grads, [p.data.numpy().copy() for p in model.parameters()]


ref = build(0)
w0 = [p.data.numpy().copy() for p in ref.parameters()]
g_ref, w_ref = one_step(ref, False)
model = build(0)
ctx = dist.init(model.parameters(), transport=FakeTransport(), bucket_mb=0.05)
g_dp, w_dp = one_step(model, True)
names = [k for k, _ in model.named_parameters()]
bad = 0
for i, (a, b, c, d) in enumerate(zip(g_dp, g_ref, w_dp, w_ref)):
    eg = np.abs(a - b).max() / max(np.abs(b).max(), 1e-12)
    ew = np.abs(c - d).max() / max(np.abs(d).max(), 1e-12)
    flag = "  <-- " if (eg > 1e-5 or ew > 1e-5) else ""
    bad += bool(flag)
    print("%2d %-28s shape %-18s grad err %.2e  update err %.2e%s" % (i, names[i] if i < len(names) else "?", c.shape, eg, ew, flag))
print("buckets:", len(ctx._plan), "bad:", bad)
