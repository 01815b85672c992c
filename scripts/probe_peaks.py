"""Measure the FP64 roofline denominators on this box (MEASURED_PEAKS.json only has HBM and bf16)."""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from matchcake_b200 import ops as mops

out = {}
out["dfma_tflops"] = mops.probe_fp64_tflops(0, 4000)
out["dmma_tflops"] = mops.probe_fp64_tflops(1, 4000)
a = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
b = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
for _ in range(2):
    a @ b
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
best = 1e9
for _ in range(3):
    e0.record(); a @ b; e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
out["cublas_dgemm_8192_tflops"] = 2 * 8192 ** 3 / (best * 1e-3) / 1e12
x = torch.empty(1 << 28, dtype=torch.float64, device="cuda"); y = torch.empty_like(x)
best = 1e9
for _ in range(5):
    e0.record(); y.copy_(x); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
out["copy_gbs"] = 2 * x.numel() * 8 / (best * 1e-3) / 1e9
out["gpu"] = torch.cuda.get_device_name(0)
print(json.dumps(out))
