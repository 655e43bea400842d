"""Data-parallel host logic (bucketing, gradient averaging, parameter broadcast) with world_size 2 and 3 on CPU:
one process per replica, torch.distributed `gloo` as the transport, the oracle's numpy device as the compute device.
Parity definition: replica r's pre-all-reduce gradients equal single-process gradients on shard
r, and the applied update equals the update from the mean of the shard gradients."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from conftest import ROOT

WORKER = textwrap.dedent('''
    import os, sys
    sys.path.insert(0, %(root)r)
    sys.path.insert(0, os.path.join(%(root)r, "tests"))
    import numpy as np
    import torch
    import torch.distributed as td
    import deepflows_b200
    from oracle import numpy_device
    from DeepFlows import backend_api, nn, tensor, dist
    from DeepFlows.tensor import Tensor
    import DeepFlows, workloads

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    td.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=rank, world_size=world)
    backend_api.register_numpy_device(numpy_device)
    backend_api.set_dgrad_mode("exact")

    class GlooTransport:
        """Same interface as DeepFlows.dist.NcclTransport, over gloo on the numpy device's host buffers."""
        def __init__(self):
            self.rank, self.world = rank, world
        def _view(self, flat):
            return flat._handle.buf[flat._offset: flat._offset + flat.size]
        def allreduce_sum(self, flat):
            t = torch.from_numpy(self._view(flat))
            td.all_reduce(t)
        def broadcast(self, flat, root=0):
            td.broadcast(torch.from_numpy(self._view(flat)), root)
        def wait(self):
            pass
        def close(self):
            pass

    df = workloads.namespace(DeepFlows)
    np.random.seed(100 + rank)                     # different init per rank: broadcast must fix it
    model = workloads.cnn_cifar10(df, "cpu", widths=(4, 8, 8), in_hw=16, dropout=0.0)
    ctx = dist.init(model.parameters(), transport=GlooTransport(), bucket_mb=0.002)   # tiny buckets: several of them
    w0 = {k: p.data.numpy().copy() for k, p in model.named_parameters()}
    rng = np.random.RandomState(7)
    X = rng.randn(4 * world, 3, 16, 16).astype(np.float32)
    T = np.eye(10, dtype=np.float32)[rng.randint(0, 10, 4 * world)]
    xs, ts = X[rank * 4:(rank + 1) * 4], T[rank * 4:(rank + 1) * 4]
    opt = df.optim.SGD(model.parameters(), lr=0.1)
    dev = backend_api.Device("cpu")
    loss = nn.CrossEntropyLoss()(model(Tensor(xs, device=dev)), Tensor(ts, device=dev))
    opt.zero_grad()
    loss.backward()
    reduced = {k: p.grad.numpy().copy() for k, p in model.named_parameters()}
    opt.step()
    w1 = {k: p.data.numpy().copy() for k, p in model.named_parameters()}
    np.savez(%(out)r + "_%%d.npz" %% rank, nbuckets=len(ctx._plan), **{"w0." + k: v for k, v in w0.items()},
             **{"g." + k: v for k, v in reduced.items()}, **{"w1." + k: v for k, v in w1.items()})
    dist.shutdown()
    td.destroy_process_group()
''')


@pytest.mark.parametrize("world", [2, 3])
def test_replicas_average_gradients(tmp_path, cpu_device, world):
    out = str(tmp_path / "rank")
    script = tmp_path / "worker.py"
    port = 29000 + (os.getpid() * 7 + world) % 1000
    script.write_text(WORKER % dict(root=ROOT, port=port, out=out))
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    logs = [p.communicate(timeout=300)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(logs)
    ranks = [dict(np.load(out + "_%d.npz" % r)) for r in range(world)]
    r0 = ranks[0]
    assert int(r0["nbuckets"]) > 1
    keys = [k[3:] for k in r0 if k.startswith("w0.")]
    for k in keys:  # broadcast: every replica starts from rank 0's weights; all-reduce: identical summed gradients
        for rr in ranks[1:]:
            assert np.array_equal(r0["w0." + k], rr["w0." + k]), k
            assert np.array_equal(r0["g." + k], rr["g." + k]), k
            assert np.array_equal(r0["w1." + k], rr["w1." + k]), k

    # single-process reference: gradient of each shard, averaged, applied with the same SGD step
    import DeepFlows
    from DeepFlows import backend_api, nn, tensor
    from DeepFlows.tensor import Tensor
    import workloads
    backend_api.set_dgrad_mode("exact")
    try:
        df = workloads.namespace(DeepFlows)
        rng = np.random.RandomState(7)
        X = rng.randn(4 * world, 3, 16, 16).astype(np.float32)
        T = np.eye(10, dtype=np.float32)[rng.randint(0, 10, 4 * world)]
        shard_grads = []
        for r in range(world):
            tensor.Graph.free_graph_all()
            model = workloads.cnn_cifar10(df, "cpu", widths=(4, 8, 8), in_hw=16, dropout=0.0)
            for k, p in model.named_parameters():
                p.data = backend_api.Btensor(r0["w0." + k], device=cpu_device)
            loss = nn.CrossEntropyLoss()(model(Tensor(X[r * 4:(r + 1) * 4], device=cpu_device)),
                                         Tensor(T[r * 4:(r + 1) * 4], device=cpu_device))
            loss.backward()
            shard_grads.append({k: p.grad.numpy().copy() for k, p in model.named_parameters()})
        for k in keys:
            total = shard_grads[0][k].copy()
            for g in shard_grads[1:]:
                total = total + g[k]
            assert np.abs(r0["g." + k] - total).max() <= 1e-6 * max(1.0, np.abs(total).max()), k
            want = r0["w0." + k] - np.float32(0.1) * (total * np.float32(1.0 / world))
            assert np.abs(r0["w1." + k] - want).max() <= 1e-6 * max(1.0, np.abs(want).max()), k
    finally:
        backend_api.set_dgrad_mode("exact")
