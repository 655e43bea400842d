"""One rank of the NCCL data-parallel parity check (launched by tests/test_gpu_dist_nccl.py, one process per GPU).

SURVEY 8e defines data-parallel parity as two identities, asserted here on the real NCCL path (bucket pack -> in-place
all-reduce on the communication stream -> gradients aliasing the buckets -> 1/world folded into the fused optimizer):
  (1) replica r's gradients BEFORE the all-reduce equal the gradients a single process computes on shard r;
  (2) the gradients the optimizer sees equal the MEAN over the shards, and after the step all replicas hold identical
      parameters, equal to a single-process step with that mean gradient.
Also: the weight broadcast at init makes differently-initialised replicas identical, and the captured-graph path (the
all-reduces become nodes of the step graph) produces the same parameters as the eager path."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import deepflows_b200  # noqa: E402,F401
import DeepFlows  # noqa: E402
import workloads  # noqa: E402
from DeepFlows import backend_api, dist, nn, tensor  # noqa: E402
from DeepFlows.tensor import Tensor  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", os.environ["RANK"]))
F32 = np.float32
dev = backend_api.cuda()
dev.set_device(local)
backend_api.set_precision("fp32")
backend_api.set_dgrad_mode("exact")
df = workloads.namespace(DeepFlows)
B = 8
rng = np.random.RandomState(7)
X = np.clip(rng.randn(B * world, 3, 32, 32), -1, 1).astype(F32)
T = (np.eye(10, dtype=F32)[rng.randint(0, 10, B * world)] * 0.95 + 0.005).astype(F32)


def build(seed):
    tensor.Graph.free_graph_all()
    np.random.seed(seed)
    return workloads.resnet_cifar(df, "cuda", widths=(8, 16, 16, 32), layers=(1, 1, 1, 1), registered=True)


def grads_on(model, xs, ts):
    loss = nn.CrossEntropyLoss()(model(Tensor(xs, device=dev)), Tensor(ts, device=dev))
    for p in model.parameters():
        p.grad = None
    loss.backward()
    return [p.grad.numpy().copy() for p in model.parameters()]


def allgather(vec):
    """Every rank's float32 vector, through the library's own all-reduce (sum of one-hot rows)."""
    table = np.zeros((world, vec.size), F32)
    table[rank] = vec
    buf = backend_api.Btensor(table.reshape(-1), device=dev)
    dev.comm_allreduce_async(buf._handle, table.size)
    dev.comm_wait()
    return buf.numpy().reshape(world, -1)


# ---- replicas start from different weights; dist.init broadcasts rank 0's -----------------------------------------------
model = build(100 + rank)
ctx = dist.init(model.parameters(), bucket_mb=0.05)          # small buckets: several of them, launched during backward
want_peer = os.environ.get("DEEPFLOWS_DP_TRANSPORT", "peer") != "nccl"
w0 = [p.data.numpy().copy() for p in model.parameters()]
head = np.concatenate([w.ravel()[:8] for w in w0])
heads = allgather(head)
assert all(np.array_equal(heads[0], heads[r]) for r in range(world)), "broadcast did not make the replicas identical"

# ---- reference quantities without data parallelism: gradients of every shard, computed locally by every rank -----------
dist_ctx_hooks = (Tensor._post_backward_hook, Tensor._grad_ready_hook)
Tensor._post_backward_hook = Tensor._grad_ready_hook = None   # plain single-process backward
single = build(0)
for p, w in zip(single.parameters(), w0):
    p.data = backend_api.Btensor(w, device=dev)
shard_grads = [grads_on(single, X[r * B:(r + 1) * B], T[r * B:(r + 1) * B]) for r in range(world)]
mean_grads = [np.mean([shard_grads[r][i] for r in range(world)], axis=0) for i in range(len(w0))]
for p, g in zip(single.parameters(), mean_grads):
    p.grad = backend_api.Btensor(g.astype(F32), device=dev)
opt1 = df.optim.SGD(single.parameters(), lr=0.1, momentum=0.9, weight_decay=5e-4)
saved_ctx, dist._ctx = dist._ctx, None                        # the reference step is a single-process one: no 1/world
opt1.step()
dist._ctx = saved_ctx
want_params = [p.data.numpy().copy() for p in single.parameters()]
Tensor._post_backward_hook, Tensor._grad_ready_hook = dist_ctx_hooks

# ---- identity (1): this replica's own gradients (hooks off so that nothing is reduced yet) ------------------------------
Tensor._post_backward_hook = Tensor._grad_ready_hook = None
mine = grads_on(model, X[rank * B:(rank + 1) * B], T[rank * B:(rank + 1) * B])
Tensor._post_backward_hook, Tensor._grad_ready_hook = dist_ctx_hooks
gmax = max(np.abs(g).max() for g in shard_grads[rank])
for i, (a, b) in enumerate(zip(mine, shard_grads[rank])):
    assert np.abs(a - b).max() <= 1e-6 * max(np.abs(b).max(), 1e-3 * gmax), "identity 1, parameter %d" % i

# ---- identity (2): the data-parallel step --------------------------------------------------------------------------------
opt = df.optim.SGD(model.parameters(), lr=0.1, momentum=0.9, weight_decay=5e-4)
loss = nn.CrossEntropyLoss()(model(Tensor(X[rank * B:(rank + 1) * B], device=dev)), Tensor(T[rank * B:(rank + 1) * B], device=dev))
opt.zero_grad()
loss.backward()                                               # buckets are reduced as they fill
scale = dist.pre_step()                                       # orders the compute stream after the reductions
assert abs(scale - 1.0 / world) < 1e-12
for i, (p, g) in enumerate(zip(model.parameters(), mean_grads)):
    got = p.grad.numpy() * scale
    assert np.abs(got - g).max() <= 2e-6 * max(np.abs(g).max(), 1e-3 * gmax), "identity 2 (mean gradient), parameter %d" % i
opt.step()
tensor.Graph.free_graph()
after = [p.data.numpy().copy() for p in model.parameters()]
for i, (a, b) in enumerate(zip(after, want_params)):
    assert np.abs(a - b).max() <= 2e-6 * max(np.abs(b).max(), 1e-6), "identity 2 (update), parameter %d" % i
sums = allgather(np.array([float(np.sum([a.astype(np.float64).sum() for a in after]))], F32))
assert all(sums[r, 0] == sums[0, 0] for r in range(world)), "replicas diverged after the step: %s" % sums[:, 0]

# ---- the captured-graph path: same parameters as two more eager steps ---------------------------------------------------
from DeepFlows.cuda_graph import CapturedStep  # noqa: E402
x_dev = Tensor(backend_api.Btensor(X[rank * B:(rank + 1) * B], device=dev))
t_dev = Tensor(backend_api.Btensor(T[rank * B:(rank + 1) * B], device=dev))


def step_fn():
    l_ = nn.CrossEntropyLoss()(model(x_dev), t_dev)
    opt.zero_grad()
    l_.backward()
    opt.step()
    tensor.Graph.free_graph()
    return l_


snapshot = [p.data.numpy().copy() for p in model.parameters()]
vel = [v.numpy().copy() for v in opt.v]
for _ in range(3):
    step_fn()
eager3 = [p.data.numpy().copy() for p in model.parameters()]
for p, w in zip(model.parameters(), snapshot):
    p.data = backend_api.Btensor(w, device=dev)   # (conv weights go back to channels-last at their next use)
for i, v in enumerate(vel):
    opt.v[i] = backend_api.Btensor(v, device=dev)
cap = CapturedStep(step_fn, device=dev, warmup=1)
for _ in range(3):
    cap()
dev.synchronize()
assert cap.captured
graph3 = [p.data.numpy().copy() for p in model.parameters()]
for i, (a, b) in enumerate(zip(graph3, eager3)):
    assert np.abs(a - b).max() <= 1e-5 * max(np.abs(b).max(), 1e-6), "captured data-parallel step differs from eager, parameter %d" % i
sums = allgather(np.array([float(np.sum([a.astype(np.float64).sum() for a in graph3]))], F32))
assert all(sums[r, 0] == sums[0, 0] for r in range(world)), "replicas diverged after the captured steps: %s" % sums[:, 0]
cap.destroy()
if want_peer:
    # the buckets really went through the peer-memory kernels (no silent NCCL fallback on a box whose GPUs see each other)
    assert getattr(ctx.transport, "peer", False), "peer-memory transport was requested but the buckets stayed on NCCL"
    ctx.transport.check()
else:
    assert not getattr(ctx.transport, "peer", False)
dist.shutdown()
print("rank %d of %d: data-parallel NCCL parity ok" % (rank, world), flush=True)
os._exit(0)
