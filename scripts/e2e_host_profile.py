import cProfile, pstats, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import matchcake_b200 as mc
from matchcake_b200.operations import CompRxRx, CompRyRy, CompRzRz, fSWAP
from matchcake_b200.observables import Projector
n, B = 16, 65536
rng = np.random.default_rng(0)
host = [torch.as_tensor(rng.uniform(0, 2 * np.pi, (B, 2))).pin_memory() for _ in range(4)]
import os
qdev = mc.NonInteractingFermionicDevice(wires=n, output_device="cpu", pinned_results=os.environ.get("PINNED", "1") == "1")
def circuit_ops(p):
    for w in range(0, n - 1, 2):
        yield CompRxRx(p, wires=[w, w + 1]); yield CompRyRy(p, wires=[w, w + 1]); yield CompRzRz(p, wires=[w, w + 1])
    for w in range(1, n - 1, 2):
        yield fSWAP(wires=[w, w + 1])
proj = Projector(np.zeros(n, dtype=int), wires=range(n))
def step(i):
    return qdev.execute_generator(circuit_ops(host[i % 4]), observable=proj, output_type="expval", reset=True)
for i in range(20): step(i)
torch.cuda.synchronize(); t = time.perf_counter()
for i in range(200): step(i)
torch.cuda.synchronize(); print("ms/step", (time.perf_counter() - t) * 5)
pr = cProfile.Profile(); pr.enable()
for i in range(200): step(i)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(25)
