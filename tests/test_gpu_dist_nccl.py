"""Data parallelism on the real NCCL path: one process per GPU on 2 GPUs of the box (skipped on a single-GPU box - the
gloo / numpy-device tests in tests/test_dist_gloo.py cover the host logic there). See tests/dp_nccl_worker.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("transport,world", [("peer", 2), ("nccl", 2), ("peer", 4)])
def test_nccl_data_parallel_identities(cuda_device, transport, world):
    """`peer`: buckets reduced by the library's own kernel over NVLink peer memory and awaited inside the optimizer kernel
    (csrc/peer.cu; the worker asserts that path was really taken); `nccl`: ncclAllReduce on the communication stream."""
    if cuda_device.device_count() < world:
        pytest.skip("needs %d GPUs (run with `gpurun --gpus %d`)" % (world, world))
    port = 29400 + (os.getpid() * 7 + world * 3 + len(transport)) % 500
    procs = []
    for rank in range(world):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   DEEPFLOWS_DP_TRANSPORT=transport)
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "dp_nccl_worker.py")], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=300)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        outs.append(out)
    for rank, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and "parity ok" in out, "rank %d:\n%s" % (rank, out[-3000:])
