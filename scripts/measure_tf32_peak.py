#!/usr/bin/env python
"""Yardstick only (never on the product path): the TF32 tensor-pipe peak of this B200 as cuBLAS reaches it through
torch.matmul with allow_tf32, the way MEASURED_PEAKS.json's bf16 figure was taken - 8192^3, best of 10 (burst) and
back to back for 3 s (sustained). Writes profiles/measured_tf32_peak.json, which bench.py reads for `roofline.peak` of
tensor-bound kernels.   python scripts/measure_tf32_peak.py"""
import json
import os
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
torch.backends.cuda.matmul.allow_tf32 = True
torch.backends.cudnn.allow_tf32 = True
n = 8192
a = torch.randn(n, n, device="cuda", dtype=torch.float32)
b = torch.randn(n, n, device="cuda", dtype=torch.float32)
for _ in range(3):
    a @ b
torch.cuda.synchronize()
best = 1e9
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    a @ b
    e1.record()
    e1.synchronize()
    best = min(best, e0.elapsed_time(e1))
burst = 2.0 * n ** 3 / (best * 1e-3) / 1e12
t0 = time.time()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
reps = 0
while time.time() - t0 < 3.0:
    for _ in range(10):
        a @ b
    reps += 10
    torch.cuda.synchronize()
e1.record()
e1.synchronize()
sustained = 2.0 * n ** 3 * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12
out = {"tf32_tflops": burst, "tf32_tflops_sustained": sustained, "gpu_name": torch.cuda.get_device_name(0), "torch": torch.__version__,
       "how": "torch.matmul fp32 8192^3 with torch.backends.cuda.matmul.allow_tf32 (cuBLAS TF32): best of 10 (burst), back to back "
              "for 3 s (sustained); CUDA events", "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime())}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "measured_tf32_peak.json"), "w"), indent=1)
print(json.dumps(out))
