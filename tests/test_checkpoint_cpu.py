"""Checkpoints on the oracle numpy device: resume == uninterrupted after real Adam steps, and the reference-written
.pkl (test/checkpoints-cifar10cuda_70%, read in place - only where the reference tree exists) loads."""
import os

import pytest

import ckpt_checks


def test_resume_equals_uninterrupted(cpu_device, tmp_path):
    ckpt_checks.resume_equals_uninterrupted("cpu", tmp_path)


@pytest.mark.skipif(not os.path.exists(ckpt_checks.REF_PKL), reason="needs the reference tree (build container)")
def test_loads_reference_written_checkpoint(cpu_device):
    ckpt_checks.loads_reference_written_checkpoint("cpu")
