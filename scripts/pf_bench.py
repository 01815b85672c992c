"""Batched Pfaffian throughput by size: python scripts/pf_bench.py [min_m]   (MCB200_PF256_SLAB=1 selects the slab
variant for 128 < m <= 256)."""
import os
import sys

import torch

sys.path.insert(0, ".")
from matchcake_b200 import ops as mops

min_m = int(sys.argv[1]) if len(sys.argv) > 1 else 0
kind = sys.argv[2] if len(sys.argv) > 2 else "gauss"   # gauss: random skew matrices; cov: Lambda_i + Lambda_j of random SPTMs
print("matrices:", kind, "| variant for 128 < m <= 256:", "slab" if os.environ.get("MCB200_PF256_SLAB") == "1" else "panel")
for m, n in [(16, 400000), (32, 200000), (64, 100000), (96, 20000), (128, 20000), (160, 8000), (192, 8000), (224, 8000),
             (256, 8000)]:
    if m < min_m:
        continue
    if kind == "cov":  # what the Gram path feeds K3: sums of two evolved vacuum covariances R^T Lambda_0 R
        ns = 256
        r = torch.linalg.qr(torch.randn(ns, m, m, dtype=torch.float64, device="cuda"))[0]
        cov = mops.cov_basis(r.contiguous(), m // 2)
        i = torch.randint(0, ns, (n,), device="cuda")
        j = (i + 1 + torch.randint(0, ns - 1, (n,), device="cuda")) % ns
        a = cov[i] + cov[j]
    else:
        a = torch.randn(n, m, m, dtype=torch.float64, device="cuda")
        a = a - a.transpose(1, 2)
    for _ in range(2):
        out = mops.pfaffian(a, signed=True)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        out = mops.pfaffian(a, signed=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    # accuracy on a small sample against log|det| / 2 (LAPACK)
    ld = torch.linalg.slogdet(a[:16])[1] / 2
    err = float(((out[:16].abs().log() - ld).abs()).max())
    print(f"m={m:4d} n={n:7d}: {ms:9.3f} ms  {n/ms/1e3:9.3f} M pf/s  {n*m**3/3/ms/1e9:7.2f} TFLOP/s  read {n*m*m*8/ms/1e6:7.0f} GB/s"
          f"  |log|Pf| - log det/2| <= {err:.1e}")
