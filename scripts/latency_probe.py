#!/usr/bin/env python
"""Where do the microseconds of a small kernel go? Graph-replayed chains (what a kernel costs inside the captured
training step) of: a trivial kernel (the launch floor with programmatic dependent launch), tcgen05 GEMMs of the
layer-3/4 tile counts at growing K (fixed cost vs cost per k-block), and BatchNorm at shrinking row counts.

    python scripts/latency_probe.py            # run again with DFB_PDL=0 for the floor without PDL
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import deepflows_b200  # noqa: E402,F401
from DeepFlows import backend_api  # noqa: E402

m = backend_api.cuda().mod
rng = np.random.RandomState(0)


def dev_rand(n):
    a = m.Array(n)
    m.from_numpy(rng.randn(n).astype(np.float32), a)
    return a


def graph_us(fn, reps=20, launches=5):
    for _ in range(3):
        fn()
    m.graph_begin_capture()
    for _ in range(reps):
        fn()
    g = m.graph_end_capture()
    for _ in range(2):
        m.graph_launch(g)
    e0, e1 = m.event_create(), m.event_create()
    m.event_record(e0)
    for _ in range(launches):
        m.graph_launch(g)
    m.event_record(e1)
    m.event_synchronize(e1)
    us = m.event_elapsed_ms(e0, e1) / (launches * reps) * 1e3
    m.event_destroy(e0)
    m.event_destroy(e1)
    m.graph_destroy(g)
    return us


print("PDL", os.environ.get("DFB_PDL", "1"))
tiny = m.Array(4)
print("fill(4 floats) chain            : %6.2f us / kernel" % graph_us(lambda: m.fill(tiny, 1.0)))
a1, b1 = dev_rand(1 << 18), m.Array(1 << 18)
print("scalar_add 1 MB chain           : %6.2f us / kernel" % graph_us(lambda: m.scalar_add(a1, 1.0, b1)))
a8, b8 = dev_rand(1 << 21), m.Array(1 << 21)
print("scalar_add 8 MB chain           : %6.2f us / kernel" % graph_us(lambda: m.scalar_add(a8, 1.0, b8)))

print("\ntcgen05 GEMM C[M,N] = A[M,K] . B[N,K]^T (TF32), per call inside a graph")
for (M, N) in [(1024, 256), (4096, 128), (16384, 64), (65536, 32)]:
    for K in (32, 128, 288, 576, 1152, 2304):
        A, B, C = dev_rand(M * K), dev_rand(N * K), m.Array(M * N)
        us = graph_us(lambda: m.gemm(A, B, C, M, N, K, 0, 1, K, K, N, 0, None, m.MODE_TF32))
        print("  M=%6d N=%4d K=%5d : %6.2f us  %7.1f TFLOP/s  %7.0f GB/s" % (M, N, K, us, 2.0 * M * N * K / us / 1e6,
                                                                             4.0 * (M * K + N * K + M * N) / us / 1e3))
        del A, B, C

print("\nconv 3x3 (TF32) fprop / dgrad / wgrad per call inside a graph")
for (n, c, h, k, s) in [(256, 32, 16, 32, 1), (256, 64, 8, 64, 1), (256, 128, 4, 128, 1), (256, 256, 2, 256, 1), (256, 128, 4, 256, 2)]:
    oh = (h + 2 - 3) // s + 1
    x, wt, y = dev_rand(n * c * h * h), dev_rand(k * c * 9), m.Array(n * oh * oh * k)
    gy, dx, dw = dev_rand(n * oh * oh * k), m.Array(n * c * h * h), m.Array(k * c * 9)
    tf = graph_us(lambda: m.conv2d_fprop(x, m.LAYOUT_NHWC, wt, y, n, c, h, h, k, 3, 1, s, m.MODE_TF32, None, 0))
    td = graph_us(lambda: m.conv2d_dgrad(gy, wt, dx, n, c, h, h, k, 3, 1, s, m.MODE_TF32, m.DGRAD_EXACT, None, 0))
    tw = graph_us(lambda: m.conv2d_wgrad(x, m.LAYOUT_NHWC, gy, dw, n, c, h, h, k, 3, 1, s, m.MODE_TF32, None, 0))
    print("  %3dx%3dx%2dx%2d -> %3d s%d : fprop %6.2f  dgrad %6.2f  wgrad %6.2f us" % (n, c, h, h, k, s, tf, td, tw))

print("\nBatchNorm forward / backward per call inside a graph")
for rows, c in [(262144, 32), (65536, 32), (16384, 64), (4096, 128), (1024, 256), (256, 256)]:
    x, gy, y, dx = dev_rand(rows * c), dev_rand(rows * c), m.Array(rows * c), m.Array(rows * c)
    g, b_, mean, inv, rm, rv, dg, db = [dev_rand(c) for _ in range(8)]
    m.fill(rv, 1.0)
    tf = graph_us(lambda: m.bn_fwd_train(x, g, b_, y, mean, inv, rm, rv, 0.1, 1e-5, rows, c))
    tb = graph_us(lambda: m.bn_bwd(x, gy, g, mean, inv, dx, dg, db, rows, c))
    print("  %7d x %3d : fwd %6.2f us  bwd %6.2f us" % (rows, c, tf, tb))
