import os, sys
import numpy as np, torch
sys.path.insert(0, ".")
from matchcake_b200 import ops as mops
d = int(sys.argv[1]); n = int(sys.argv[2])
torch.manual_seed(0)
q, _ = torch.linalg.qr(torch.randn(n, d, d, dtype=torch.float64, device="cuda"))
j0 = torch.zeros(d, d, dtype=torch.float64, device="cuda")
for k in range(d // 2): j0[2*k, 2*k+1] = -1; j0[2*k+1, 2*k] = 1
cov = q.transpose(1, 2) @ j0 @ q
cov = (cov - cov.transpose(1, 2)) / 2
out = mops.gram(cov, cov, 1.0, mode=mops.GRAM_FULL)
ref = torch.sqrt(torch.abs(torch.linalg.det(cov[:, None] + cov[None, :])))
err = (out - ref).abs() / (ref.abs() + 1e-300)
print(os.environ.get("MCB200_GRAM_TILES", "rows"), d, n, "max rel err", float(err.max()), "n bad (>1e-8)", int((err > 1e-8).sum()), "of", n * n)
bad = torch.nonzero(err > 1e-8)[:5]
for i, j in bad.tolist(): print(i, j, float(out[i, j]), float(ref[i, j]))
