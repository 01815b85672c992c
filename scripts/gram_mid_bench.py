import sys, torch
sys.path.insert(0, ".")
from matchcake_b200 import ops as mops
for d, n in [(96, 512), (128, 512)]:
    a = torch.randn(n, d, d, dtype=torch.float64, device="cuda"); a = (a - a.transpose(1, 2)) / d ** 0.5
    for _ in range(2): k = mops.gram(a, a, 1.0)
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); k = mops.gram(a, a, 1.0); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1); cells = n * (n - 1) // 2
    print(f"d={d} n={n}: {ms:.2f} ms  {cells/ms/1e3:.2f} M entries/s  {cells*d**3/3/ms/1e9:.2f} TFLOP/s")
