#!/usr/bin/env python
"""TF32 tcgen05 GEMM throughput for the four operand-major combinations (M=N=K=4096 and a wgrad-like skinny shape)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import deepflows_b200  # noqa
from DeepFlows import backend_api
m = backend_api.cuda().mod
rng = np.random.RandomState(0)
def dev_rand(n):
    a = m.Array(n); m.from_numpy(rng.randn(n).astype(np.float32), a); return a
for (M, N, K) in [(4096, 4096, 4096), (128, 128, 65536), (32, 32, 65536), (8192, 128, 1152)]:
    for ta in (0, 1):
        for tb in (0, 1):
            A, B, C = dev_rand(M * K), dev_rand(K * N), m.Array(M * N)
            lda, ldb = (M if ta else K), (K if tb else N)
            f = lambda: m.gemm(A, B, C, M, N, K, ta, tb, lda, ldb, N, 0, None, m.MODE_TF32)
            for _ in range(3): f()
            e0, e1 = m.event_create(), m.event_create()
            m.event_record(e0)
            for _ in range(10): f()
            m.event_record(e1); m.event_synchronize(e1)
            us = m.event_elapsed_ms(e0, e1) / 10 * 1e3
            print("M=%d N=%d K=%d ta=%d tb=%d : %8.1f us  %7.1f TFLOP/s" % (M, N, K, ta, tb, us, 2.0 * M * N * K / us / 1e6))
