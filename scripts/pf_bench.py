import sys, torch
sys.path.insert(0, ".")
from matchcake_b200 import ops as mops
for m, n in [(16, 400000), (32, 200000), (64, 100000), (96, 20000), (128, 20000), (192, 4000), (256, 4000)]:
    a = torch.randn(n, m, m, dtype=torch.float64, device="cuda"); a = a - a.transpose(1, 2)
    for _ in range(2): mops.pfaffian(a)
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): mops.pfaffian(a)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"m={m:4d} n={n:7d}: {ms:9.3f} ms  {n/ms/1e3:9.2f} M pf/s  {n*m**3/3/ms/1e9:7.2f} TFLOP/s  read {n*m*m*8/ms/1e6:7.0f} GB/s")
