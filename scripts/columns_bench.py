import sys, torch, numpy as np
sys.path.insert(0, ".")
from matchcake_b200 import ops as mops
def brick(n, depth):
    gates, off = [], 0
    kinds = (mops.GATE_RXRX, mops.GATE_RYRY, mops.GATE_RZRZ)
    for l in range(depth):
        for w in range(l % 2, n - 1, 2):
            gates.append(mops.GateSpec(kinds[l % 3], w, w + 1, off)); off += 2
    return gates, off
for n, depth, B in [(64, 64, 4096), (32, 32, 8192), (16, 16, 16384)]:
    gates, off = brick(n, depth)
    circ = mops.CompiledCircuit(n, gates, off)
    params = torch.rand((B, off), dtype=torch.float64, device="cuda") * 6.28
    for m in (2, 4, 8, 16, 32):
        if m > 2 * n: continue
        cols = torch.arange(m, dtype=torch.int32, device="cuda")
        res = {}
        for name, flag in (("brick", 1), ("smem", 0)):
            circ._struct.nearest_neighbour = flag
            for _ in range(2): x = mops.circuit_columns(circ, params, cols)
            torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3): x = mops.circuit_columns(circ, params, cols)
            e1.record(); torch.cuda.synchronize()
            res[name] = (e0.elapsed_time(e1) / 3, x)
        err = (res["brick"][1] - res["smem"][1]).abs().max().item()
        print(f"N={n} depth={depth} B={B} m={m}: brick {res['brick'][0]:.3f} ms  smem {res['smem'][0]:.3f} ms  maxdiff {err:.1e}")
