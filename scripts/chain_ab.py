"""Timings that depend on warp_pfaffian_static_steps: C2 fused projector kernel, plain Pfaffians (m = 32, 64), outcome
probabilities (C4's read-out shape), 64 x 64 Gram (old kernel via MCB200_GRAM_FLAT=0).  Run once per library build:
MCB200_LIB=matchcake_b200/libmcb200_chain_old.so python scripts/chain_ab.py
The comparison build: MCB200_NVCC_EXTRA="-DMCB200_STATIC_CHAIN_OLD" python -m matchcake_b200.build --force, copy the .so
aside, rebuild without the flag.  Measured (one B200): C2 170.8 -> 169.4 us, Pfaffians m = 32 906 -> 846 us per 262144,
m = 64 1575 -> 1603 us per 65536, 2048-sample Gram with the branchy kernel 15.85 -> 15.72 ms."""
import os, sys
import numpy as np, torch
sys.path.insert(0, ".")
from matchcake_b200 import ops as mops
from bench import readme_gate_specs


def timed(fn, reps=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


rng = np.random.default_rng(0)
print("lib", os.environ.get("MCB200_LIB", "default"))
n, B = 16, 65536
circ = mops.CompiledCircuit(n, readme_gate_specs(n, mops), 2)
pool = [torch.as_tensor(rng.uniform(0, 2 * np.pi, (B, 2))).cuda() for _ in range(64)]
it = iter(range(10 ** 9))
out = mops.circuit_expval_projector(circ, pool[0])
print(f"c2 expval: {timed(lambda: mops.circuit_expval_projector(circ, pool[next(it) % 64]), reps=100) * 1e3:.1f} us  checksum {float(out.sum()):.12e}")
for m, cnt in ((32, 1 << 18), (64, 1 << 16)):
    a = torch.randn(cnt, m, m, dtype=torch.float64, device="cuda")
    a = a - a.transpose(1, 2)
    r = mops.pfaffian(a, signed=True)
    print(f"pfaffian m={m}: {timed(lambda: mops.pfaffian(a, signed=True)) * 1e3:.1f} us for {cnt}  checksum {float(r.abs().log().sum()):.12e}")
os.environ["MCB200_GRAM_FLAT"] = "0"
q, _ = torch.linalg.qr(torch.randn(2048, 64, 64, dtype=torch.float64, device="cuda"))
j0 = torch.zeros(64, 64, dtype=torch.float64, device="cuda")
for k in range(32):
    j0[2 * k, 2 * k + 1], j0[2 * k + 1, 2 * k] = 1.0, -1.0
cov = q @ j0 @ q.transpose(1, 2)
g = mops.gram(cov, cov, 2.0 ** -32)
print(f"gram 2048 x 64 (branchy kernel): {timed(lambda: mops.gram(cov, cov, 2.0 ** -32), reps=5, warm=2):.2f} ms  checksum {float(g.sum()):.12e}")
