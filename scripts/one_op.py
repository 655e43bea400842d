#!/usr/bin/env python
"""Run one conv geometry (fprop, dgrad, wgrad) a few times - the target for `ncu --set full -k regex:...`.
    python scripts/one_op.py N C H W K R pad stride [reps]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import deepflows_b200  # noqa: E402,F401
from DeepFlows import backend_api  # noqa: E402

n, c, h, w, k, r, p, s = [int(v) for v in sys.argv[1:9]]
reps = int(sys.argv[9]) if len(sys.argv) > 9 else 3
m = backend_api.cuda().mod
rng = np.random.RandomState(0)
oh, ow = (h + 2 * p - r) // s + 1, (w + 2 * p - r) // s + 1


def dev_rand(cnt):
    a = m.Array(cnt)
    m.from_numpy(rng.randn(cnt).astype(np.float32), a)
    return a


x, wt, y = dev_rand(n * c * h * w), dev_rand(k * c * r * r), m.Array(n * oh * ow * k)
gy, dx, dw = dev_rand(n * oh * ow * k), m.Array(n * c * h * w), m.Array(k * c * r * r)
for _ in range(reps):
    m.conv2d_fprop(x, m.LAYOUT_NHWC, wt, y, n, c, h, w, k, r, p, s, m.MODE_TF32, None, 0)
    m.conv2d_dgrad(gy, wt, dx, n, c, h, w, k, r, p, s, m.MODE_TF32, m.DGRAD_EXACT, None, 0)
    m.conv2d_wgrad(x, m.LAYOUT_NHWC, gy, dw, n, c, h, w, k, r, p, s, m.MODE_TF32, None, 0)
m.synchronize()
print("ok")
