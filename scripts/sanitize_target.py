"""One small call of every kernel family, for compute-sanitizer (memcheck / racecheck):
   compute-sanitizer --tool memcheck python scripts/sanitize_target.py"""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from bench import readme_gate_specs
from matchcake_b200 import ops as mops
import matchcake_b200 as mc
from matchcake_b200.ml.kernels import FermionicPQCKernel

dev = torch.device("cuda")
rng = np.random.default_rng(0)
def skew(n, m):
    a = torch.randn(n, m, m, dtype=torch.float64, device=dev)
    return a - a.transpose(1, 2)
# K3 all size classes, static + pivoted (a zero natural pivot forces the hand-over)
for m in (2, 6, 8, 16, 24, 30, 32, 48, 64, 66, 96, 128, 130, 200, 256, 258):
    a = skew(3, m); a[1, 0, 1] = a[1, 1, 0] = 0.0
    if m >= 8: a[2, 4, 5] = a[2, 5, 4] = 0.0
    mops.pfaffian(a, signed=True)
# fused expval, all tile sizes; chain kernels (general, brick), conjugation (kernel / gemm), probs, probs_all, sampling
for n in (2, 4, 7, 12, 16, 24, 32):
    circ = mops.CompiledCircuit(n, readme_gate_specs(n, mops), 2)
    p = torch.rand(5, 2, dtype=torch.float64, device=dev) * 6.28
    mops.circuit_expval_projector(circ, p)
for n in (5, 16, 40, 64, 70):
    circ = mops.CompiledCircuit(n, readme_gate_specs(n, mops), 2)
    p = torch.rand(3, 2, dtype=torch.float64, device=dev) * 6.28
    x = mops.circuit_columns(circ, p)
    cov = mops.cov_basis(x, n)
    k = min(n, 5)
    cols = torch.arange(2 * k, dtype=torch.int32, device=dev)
    xs = mops.circuit_columns(circ, p, cols)
    cs = mops.cov_basis(xs, n)
    mops.probs_all(cs, normalize=True)
    mops.probs(cs, torch.randint(0, 2, (7, k), dtype=torch.int8, device=dev))
    mops.sample(cs, 9, seed=1, return_prob=True)
    mops.probs(cov, torch.zeros((1, n), dtype=torch.int8, device=dev))
    circ._struct.nearest_neighbour = 0
    mops.circuit_columns(circ, p, cols)
# Gram: tiles (d <= 64), block tiles (<= 128), slab (<= 256), block-factorised
for d, ns in ((8, 20), (32, 12), (64, 10), (96, 5), (160, 3)):
    c = skew(ns, d) / d ** 0.5
    mops.gram(c, c, 1.0)
    mops.gram(c[:3], c, 1.0, mode=mops.GRAM_FULL)
b0 = torch.randn(70, 5, 6, dtype=torch.float64, device=dev); b1 = torch.randn(130, 5, 6, dtype=torch.float64, device=dev)
mops.gram_blocks4(b0, b1); mops.gram_blocks4(b1, b1)
kern = FermionicPQCKernel(n_qubits=6, rotations="X,Z", r_dtype=torch.float64, c_dtype=torch.complex128)
xx = rng.random((9, 6)); kern.fit(xx); kern(xx)
# dense chain (TMA + tiled), dense conjugation, prep kernels, Pauli sums
for n in (6, 16, 32, 64, 70):
    a = torch.randn(3, n, n, dtype=torch.float64, device=dev)
    mops.sptm_matmul(a, a, False, True); mops.sptm_matmul(a, a[0], True, False)
r = torch.tensor(np.stack([np.eye(8)] * 2), device=dev)
amp = torch.randn(4, 2, dtype=torch.complex128, device=dev); amp = amp / amp.abs().pow(2).sum(-1, keepdim=True).sqrt()
l0 = mops.product_state_cov(amp, extended=True)
ext = mops.cov_dense(r, l0)
mops.sector_features(ext, torch.tensor([[0, 1, 2, 3], [2, 5, 6, 8]], dtype=torch.int32, device=dev))
mops.pauli_sum_small(ext, torch.tensor([0, 0, 2, 6, 12], dtype=torch.int32), torch.tensor([0, 1, 0, 1, 2, 3, 0, 2, 3, 4, 6, 8], dtype=torch.int32),
                     torch.ones(4, dtype=torch.float64))
mops.matchgate_sptm(torch.eye(4, dtype=torch.complex128, device=dev)[None].repeat(3, 1, 1))
# reverse mode
circ = mops.CompiledCircuit(5, readme_gate_specs(5, mops), 2)
pg = (torch.rand(4, 2, dtype=torch.float64, device=dev) * 6.28).requires_grad_()
mops.circuit_expval_projector(circ, pg).sum().backward()
a = skew(3, 70).requires_grad_(); mops.pfaffian(a).sum().backward()
rg = torch.tensor(np.stack([np.eye(8)] * 2), device=dev).requires_grad_()
mops.cov_dense(rg, l0).sum().backward()
torch.cuda.synchronize()
print("sanitize target done")
