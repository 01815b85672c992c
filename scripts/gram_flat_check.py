"""A/B of the C3-shaped Gram kernels (d = 64): branch-free steps (default) vs one cell at a time (MCB200_GRAM_FLAT=0).
Covariances of the C3 ansatz on n samples; parity of the two against each other and against sqrt|det| on a sample."""
import os, sys
import numpy as np, torch
sys.path.insert(0, ".")
from matchcake_b200 import ops as mops
from matchcake_b200.ml.kernels import FermionicPQCKernel

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
nq = 32
x = np.random.default_rng(0).random((n, 2 * nq))
kern = FermionicPQCKernel(n_qubits=nq, rotations="X,Z", r_dtype=torch.float64, c_dtype=torch.complex128)
kern.float32_gram = False
kern.fit(x)
cov = kern._x_to_cov(torch.as_tensor(x, device="cuda"))
scale = 2.0 ** (-nq)
res = {}
for mode in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["0", "1", "0", "1"]):
    os.environ["MCB200_GRAM_FLAT"] = mode
    for _ in range(2):
        k = mops.gram(cov, cov, scale)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); k = mops.gram(cov, cov, scale); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1); cells = n * (n - 1) // 2
    print(f"flat={mode} n={n}: {ms:.2f} ms  {cells / ms / 1e3:.2f} M entries/s  "
          f"{cells * (64 ** 3 / 3 + 64 ** 2) / ms / 1e9:.2f} TFLOP/s", flush=True)
    res[mode] = k.clone()
a = res["0"]
for key, b in res.items():
    print(key, "vs plain: max rel diff", ((a - b).abs() / (a.abs() + 1e-300)).max().item())
b = res[list(res)[-1]]
rel = ((a - b).abs() / (a.abs() + 1e-300)).max().item()
print("flat vs branchy: max rel diff", rel, "max abs diff", (a - b).abs().max().item())
# the default (flat) kernel against sqrt|det(L_i + L_j)| on a sample of cells (rows 0..63 against everything)
m = min(64, n)
ref = scale * torch.sqrt(torch.abs(torch.linalg.det(cov[:m, None] + cov[None, :])))
got = b[:m]
mask = ~torch.eye(n, dtype=torch.bool, device="cuda")[:m]
err = ((got - ref).abs() / (ref.abs() + 1e-300))[mask]
print("flat vs det: max rel", err.max().item())
# mirror / diagonal conventions identical
print("diag equal", torch.equal(a.diagonal(), b.diagonal()), "symmetric", torch.equal(b, b.T))
assert rel < 1e-11 and err.max().item() < 1e-9
print("OK")
