#!/usr/bin/env python
"""Data-parallel smoke test, one process per GPU (torchrun): NCCL id exchange, communicator, all-reduce,
broadcast, then two training steps of a small ResNet eagerly and through a captured CUDA graph. Prints a
marker after every stage (flush) so that a hang can be located."""
import os, sys, time, faulthandler
faulthandler.dump_traceback_later(20, exit=True)
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import deepflows_b200  # noqa
import workloads
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
def mark(msg):
    print("[rank %d %.1fs] %s" % (rank, time.time() - T0, msg), flush=True)
T0 = time.time()
import DeepFlows
from DeepFlows import backend_api, dist, nn
from DeepFlows.tensor import Tensor, Graph
from DeepFlows.cuda_graph import CapturedStep
dev = backend_api.cuda()
dev.set_device(local)
mark("device %d of %d: %s" % (dev.get_device(), dev.device_count(), dev.device_info()))
df = workloads.namespace(DeepFlows)
backend_api.set_precision("tf32"); backend_api.set_dgrad_mode("exact")
np.random.seed(0)
model = workloads.resnet_cifar(df, "cuda", widths=(8, 16, 16, 32), layers=(1, 1, 1, 1), registered=True)
opt = df.optim.Adam(model.parameters(), lr=1e-3)
crit = nn.CrossEntropyLoss()
mark("model built")
ctx = dist.init(model.parameters())
mark("dist.init done (world %d)" % ctx.world)
v = backend_api.Btensor(np.full(1024, rank + 1.0, np.float32), device=dev)
dev.comm_allreduce_async(v._handle, 1024); dev.comm_wait()
got = v.numpy()
mark("allreduce -> %s (want %s)" % (got[0], sum(range(1, world + 1))))
rng = np.random.RandomState(100 + rank)
x = Tensor(backend_api.Btensor(np.clip(rng.randn(16, 3, 32, 32), -1, 1).astype(np.float32), device=dev))
t = Tensor(backend_api.Btensor((np.eye(10, dtype=np.float32)[rng.randint(0, 10, 16)]), device=dev))
def step():
    loss = crit(model(x), t)
    opt.zero_grad(); loss.backward(); opt.step()
    Graph.free_graph()
    return loss
for i in range(2):
    l = step(); mark("eager step %d loss %.5f" % (i, float(l.data.numpy()[0])))
cap = CapturedStep(step, device=dev, warmup=1)
for i in range(4):
    l = cap(); dev.synchronize(); mark("graph-mode call %d (captured=%s) loss %.5f" % (i, cap.captured, float(l.data.numpy()[0])))
p0 = [p.data.numpy().ravel()[:4] for p in model.parameters()][:2]
mark("params head %s" % (p0,))
dist.shutdown()
mark("done")
