import sys, torch, numpy as np
sys.path.insert(0, ".")
from bench import readme_gate_specs
from matchcake_b200 import ops as mops
for n, B in [(40, 4096), (48, 4096), (59, 4096), (64, 4096)]:
    circ = mops.CompiledCircuit(n, readme_gate_specs(n, mops), 2)
    p = torch.rand(B, 2, dtype=torch.float64, device="cuda") * 6.28
    tgt = torch.zeros(1, n, dtype=torch.int8, device="cuda")
    def fused(): return mops.circuit_expval_projector(circ, p)
    def comp(): return mops.probs(mops.cov_basis(mops.circuit_columns(circ, p), n), tgt)[:, 0]
    for name, f in (("fused", fused), ("composite", comp)):
        if name == "fused" and n > 59: continue
        for _ in range(2): r = f()
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); r = f(); e1.record(); torch.cuda.synchronize()
        print(f"N={n} B={B} {name}: {e0.elapsed_time(e1):.2f} ms  {float(r.sum()):.6e}")
