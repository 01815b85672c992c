"""Throughput of the SURVEY 8(f) components: sampling, reverse mode, Hamiltonian read-out."""
import sys, time, torch, numpy as np
sys.path.insert(0, ".")
from bench import readme_gate_specs
from matchcake_b200 import ops as mops
def timed(f, reps=5):
    for _ in range(2): f()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
# sampling: N = 16 README circuit, 4096 instances x 1024 shots
n, B, shots = 16, 4096, 1024
circ = mops.CompiledCircuit(n, readme_gate_specs(n, mops), 2)
p = torch.rand(B, 2, dtype=torch.float64, device="cuda") * 6.28
cov = mops.cov_basis(mops.circuit_columns(circ, p), n)
ms = timed(lambda: mops.sample(cov, shots, seed=1))
print(f"sampling N={n}: {B*shots/ms/1e3:.1f} M bit strings/s ({ms:.2f} ms for {B}x{shots})")
n2 = 32
circ2 = mops.CompiledCircuit(n2, readme_gate_specs(n2, mops), 2)
cov2 = mops.cov_basis(mops.circuit_columns(circ2, p), n2)
ms = timed(lambda: mops.sample(cov2, 256, seed=1))
print(f"sampling N={n2}: {B*256/ms/1e3:.1f} M bit strings/s ({ms:.2f} ms for {B}x256)")
# reverse mode: C2 batch, d expval / d params
B = 65536
pg = (torch.rand(B, 2, dtype=torch.float64, device="cuda") * 6.28).requires_grad_()
def fwd_bwd():
    pg.grad = None
    mops.circuit_expval_projector(circ, pg).sum().backward()
ms_fb = timed(fwd_bwd, 3)
ms_f = timed(lambda: mops.circuit_expval_projector(circ, pg.detach()), 10)
print(f"C2 batch {B}: forward (fused, inference) {ms_f:.3f} ms; forward + backward (K1->K2->K3 + adjoints) {ms_fb:.2f} ms "
      f"= {B/ms_fb/1e3:.2f} M gradients/s")
# Hamiltonian read-out: nearest-neighbour XX + YY + ZZ + Z on N = 16, batch 65536
from matchcake_b200.utils.jordan_wigner import JordanWigner
from matchcake_b200.device import _PauliStructureCache
words = []
for w in range(n - 1):
    for a in "XYZ":
        s = ["I"] * n; s[w] = a; s[w + 1] = a; words.append("".join(s))
for w in range(n):
    s = ["I"] * n; s[w] = "Z"; words.append("".join(s))
structure = _PauliStructureCache.get(tuple(words), n)
sizes, flat, scal = [], [], []
for sector, (idx, ph, terms, wick) in structure.items():
    sizes += [sector] * len(terms); flat.append(idx.reshape(-1)); scal.append(np.real(ph * wick))
tp = torch.as_tensor(np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)).cuda()
fl = torch.as_tensor(np.concatenate(flat).astype(np.int32)).cuda(); sc = torch.as_tensor(np.concatenate(scal)).cuda()
covb = mops.cov_basis(mops.circuit_columns(circ, pg.detach()), n)
ext = torch.nn.functional.pad(covb, (0, 1, 0, 1)).contiguous()
ms = timed(lambda: mops.pauli_sum_small(ext, tp, fl, sc), 10)
print(f"Hamiltonian read-out ({len(words)} Pauli terms, sectors {sorted(set(sizes))}), batch {B}: {ms:.3f} ms = {B/ms/1e3:.1f} M expvals/s")
