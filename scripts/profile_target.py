"""Small single-GPU targets for ncu: python scripts/profile_target.py {c2|gram|c4|c4cols|blocks|matmul|pf|pfcov} [size]"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from bench import readme_gate_specs
from matchcake_b200 import ops as mops

what = sys.argv[1]
size = int(sys.argv[2]) if len(sys.argv) > 2 else 0
dev = torch.device("cuda")
rng = np.random.default_rng(0)
if what == "c2":
    n, B = 16, size or 65536
    circ = mops.CompiledCircuit(n, readme_gate_specs(n, mops), 2)
    p = torch.as_tensor(rng.uniform(0, 2 * np.pi, (B, 2))).to(dev)
    for _ in range(3):
        out = mops.circuit_expval_projector(circ, p)
elif what == "gram":
    from matchcake_b200.ml.kernels import FermionicPQCKernel

    n, ns = 32, size or 1024
    x = rng.random((ns, 2 * n))
    k = FermionicPQCKernel(n_qubits=n, rotations="X,Z", r_dtype=torch.float64, c_dtype=torch.complex128)
    k.fit(x)
    for _ in range(2):
        out = k(torch.as_tensor(x).to(dev))
elif what == "c4":
    n, B, depth, k = 64, size or 4096, 64, 8
    gates = []; off = 0
    kinds = (mops.GATE_RXRX, mops.GATE_RYRY, mops.GATE_RZRZ)
    for l in range(depth):
        for w in range(l % 2, n - 1, 2):
            gates.append(mops.GateSpec(kinds[l % 3], w, w + 1, off)); off += 2
    circ = mops.CompiledCircuit(n, gates, off)
    params = torch.rand((B, off), dtype=torch.float64, device=dev) * (2 * np.pi)
    maj = torch.arange(2 * k, dtype=torch.int32, device=dev)
    for _ in range(2):
        out = mops.probs_all(mops.circuit_cov_basis(circ, params, maj), normalize=True)   # fused K1 + K2, then the tree
elif what == "c4cols":  # K1 alone on C4's 16-column panel (row-major kernel, light-cone pruned gate list)
    n, B, depth, k = 64, size or 4096, 64, 8
    gates = []; off = 0
    kinds = (mops.GATE_RXRX, mops.GATE_RYRY, mops.GATE_RZRZ)
    for l in range(depth):
        for w in range(l % 2, n - 1, 2):
            gates.append(mops.GateSpec(kinds[l % 3], w, w + 1, off)); off += 2
    circ = mops.CompiledCircuit(n, gates, off)
    params = torch.rand((B, off), dtype=torch.float64, device=dev) * (2 * np.pi)
    for _ in range(2):
        out = mops.circuit_columns(circ, params, np.arange(2 * k))
elif what == "blocks":
    n = size or 32768
    b0 = torch.randn(n, 64, 6, dtype=torch.float64, device=dev)
    b1 = torch.randn(4096, 64, 6, dtype=torch.float64, device=dev)
    for _ in range(2):
        out = mops.gram_blocks4(b0, b1)
elif what == "matmul":
    # dense SPTM chaining (K1 dense): sptm_matmul_tma_kernel for n <= 64, bgemm_dmma_kernel above
    n = size or 64
    B = max(256, (1 << 26) // (n * n))  # 512 MB per operand: larger than L2
    a = torch.randn(B, n, n, dtype=torch.float64, device=dev)
    b = torch.randn_like(a)
    for _ in range(3):
        out = mops.sptm_matmul(a, b)
elif what == "pfcov":  # Pfaffians of Lambda_i + Lambda_j (what the Gram path feeds K3)
    m = size or 256
    ns, n = 256, 8000
    r = torch.linalg.qr(torch.randn(ns, m, m, dtype=torch.float64, device=dev))[0]
    cov = mops.cov_basis(r.contiguous(), m // 2)
    i = torch.randint(0, ns, (n,), device=dev)
    j = (i + 1 + torch.randint(0, ns - 1, (n,), device=dev)) % ns
    a = cov[i] + cov[j]
    for _ in range(3):
        out = mops.pfaffian(a)
elif what == "pf":
    m = size or 64
    a = torch.randn(20000, m, m, dtype=torch.float64, device=dev)
    a = a - a.transpose(1, 2)
    for _ in range(3):
        out = mops.pfaffian(a)
torch.cuda.synchronize()
print("done", float(out.sum()))
