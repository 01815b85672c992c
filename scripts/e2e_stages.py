import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import matchcake_b200 as mc
from matchcake_b200 import ops as mops
from matchcake_b200.operations import CompRxRx, CompRyRy, CompRzRz, fSWAP
from matchcake_b200.observables import Projector
n, B = 16, 65536
rng = np.random.default_rng(0)
host = [torch.as_tensor(rng.uniform(0, 2 * np.pi, (B, 2))).pin_memory() for _ in range(4)]
qdev = mc.NonInteractingFermionicDevice(wires=n, output_device="cpu")
def circuit_ops(p):
    out = []
    for w in range(0, n - 1, 2):
        out += [CompRxRx(p, wires=[w, w + 1]), CompRyRy(p, wires=[w, w + 1]), CompRzRz(p, wires=[w, w + 1])]
    for w in range(1, n - 1, 2):
        out.append(fSWAP(wires=[w, w + 1]))
    return out
proj = Projector(np.zeros(n, dtype=int), wires=range(n))
def timeit(f, reps=200):
    for _ in range(10): f()
    torch.cuda.synchronize(); t = time.perf_counter()
    for i in range(reps): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t) / reps * 1e3
i = [0]
def full():
    i[0] += 1
    return qdev.execute_generator(circuit_ops(host[i[0] % 4]), observable=proj, output_type="expval", reset=True)
print("full e2e ms", timeit(full))
print("op construction ms", timeit(lambda: circuit_ops(host[0])))
ops_ = circuit_ops(host[0])
def apply_only():
    qdev.execute_generator(ops_, reset=True)
print("apply (no read-out) ms", timeit(apply_only))
def apply_compile():
    qdev.execute_generator(ops_, reset=True); qdev._compile()
print("apply + compile (H2D) ms", timeit(apply_compile))
pd = host[0].cuda()
circ = qdev._compile()[0][0][1]
print("raw kernel via ops ms", timeit(lambda: mops.circuit_expval_projector(circ, pd)))
print("H2D only ms", timeit(lambda: host[0].to("cuda", non_blocking=True)))
out = mops.circuit_expval_projector(circ, pd)
print("D2H only ms", timeit(lambda: out.to("cpu")))
def h2d_kernel_d2h():
    p = host[0].to("cuda", non_blocking=True)
    return mops.circuit_expval_projector(circ, p).to("cpu")
print("H2D + kernel + D2H ms", timeit(h2d_kernel_d2h))
import cProfile, pstats
qdev.execute_generator(ops_, reset=True)
pr = cProfile.Profile(); pr.enable()
for _ in range(200):
    qdev.execute_output(observable=proj, output_type="expval")
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
