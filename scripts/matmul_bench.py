import torch, time
from matchcake_b200 import ops as mops
for n, B in [(16, 65536), (32, 32768), (64, 8192), (128, 2048)]:
    a = torch.randn(B, n, n, dtype=torch.float64, device="cuda"); b = torch.randn_like(a)
    for f, name in [(lambda: mops.sptm_matmul(a, b), "mcb200"), (lambda: torch.bmm(a, b), "cublas")]:
        for _ in range(3): f()
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): f()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"n={n} B={B} {name}: {ms:.3f} ms  {2*n**3*B/ms/1e9:.2f} TFLOP/s  {3*n*n*8*B/ms/1e6:.0f} GB/s")
