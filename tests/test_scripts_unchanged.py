"""north_star: "every `device='cuda'` script in test/ runs unchanged". tests/script_runner.py executes the reference's own
script source (read from $DEEPFLOWS_REFERENCE/test, default /root/reference/test - never copied here) with stubbed plotting /
datasets and a stop after a few optimizer steps, and records the loss of every step.

CPU tier (build container): each script runs once against the REFERENCE package (its host code on its numpy device with the
CUDA-semantics setitem = the oracle's definition) and once against this repo's host package (oracle numpy device
standing in for 'cuda'): same seed, same synthetic data - the step losses must agree. This pins the whole host layer (module
registries, init RNG order, data loader, schedulers, fused-op graph) against the reference, script by script.
GPU tier: the same scripts on libdfb200 against the oracle run of this package. The GPU box has no reference tree, so the
GPU tier only runs where DEEPFLOWS_REFERENCE points at one (the build container has no GPU): it is skipped with that reason
by the driver; tests/test_gpu_train.py / test_gpu_fullsize.py run the same models written out in workloads.py there."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("DEEPFLOWS_REFERENCE", "/root/reference")
RUNNER = os.path.join(ROOT, "tests", "script_runner.py")
# script, optimizer steps, synthetic samples
SCRIPTS = [("MLP_MNIST_cuda.py", 3, 40), ("CNN_MNIST_cuda.py", 2, 192), ("CNN_CIFAR10_cuda.py", 2, 40), ("ResNet_CIFAR10_cuda.py", 2, 40)]
needs_reference = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "test")),
                                     reason="the reference's scripts live in $DEEPFLOWS_REFERENCE/test (absent on the GPU box)")


def _run(script, package, device, steps, samples, tmp_path):
    out = str(tmp_path / ("%s_%s_%s.npz" % (script, package, device)))
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1")
    p = subprocess.run([sys.executable, "-B", RUNNER, "--script", script, "--package", package, "--device", device, "--steps", str(steps),
                        "--samples", str(samples), "--out", out], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env,
                       timeout=900)
    assert p.returncode == 0, p.stdout[-3000:]
    return np.load(out)["losses"]


@needs_reference
@pytest.mark.parametrize("script,steps,samples", SCRIPTS)
def test_script_runs_unchanged_like_the_reference(script, steps, samples, tmp_path):
    want = _run(script, "reference", "oracle", steps, samples, tmp_path)
    got = _run(script, "ours", "oracle", steps, samples, tmp_path)
    assert len(want) == steps and len(got) == steps
    assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max(), (got, want)


@pytest.mark.gpu
@needs_reference
@pytest.mark.parametrize("script,steps,samples", SCRIPTS)
def test_script_runs_unchanged_on_cuda(script, steps, samples, tmp_path):
    want = _run(script, "ours", "oracle", steps, samples, tmp_path)
    got = _run(script, "ours", "cuda", steps, samples, tmp_path)
    assert np.abs(got - want).max() <= 1e-4 * np.abs(want).max(), (got, want)
