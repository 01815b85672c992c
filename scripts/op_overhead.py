"""Host cost of one call through torch.ops.mcb200 vs the bare ctypes launch (tiny problem, launch-bound)."""
import sys
import time

import torch

sys.path.insert(0, ".")

from matchcake_b200 import _lib, ops as mops

a = torch.randn(4, 4, 4, dtype=torch.float64, device="cuda")
a = a - a.transpose(1, 2)
out = torch.empty(4, dtype=torch.float64, device="cuda")
lib = _lib.load()
st = torch.cuda.current_stream().cuda_stream


def timed(f, n=20000):
    for _ in range(200):
        f()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(n):
        f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / n * 1e6


print(f"ctypes mcb200_pfaffian            {timed(lambda: lib.mcb200_pfaffian(a.data_ptr(), 4, 4, 0, 1.0, out.data_ptr(), None, st)):7.2f} us/call")
print(f"torch.ops.mcb200.pfaffian         {timed(lambda: torch.ops.mcb200.pfaffian(a, False, 1.0, 0.0)):7.2f} us/call")
print(f"matchcake_b200.ops.pfaffian       {timed(lambda: mops.pfaffian(a)):7.2f} us/call")
