#!/usr/bin/env python
"""Times both routes of the fused projector kernel on config C2 (N=16 README circuit, batch 65536) and sweeps the
launch shape of the covariance route (MCB200_COV_IB / MCB200_COV_BPS).  One B200, CUDA events, rotating inputs."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from matchcake_b200 import _lib
from matchcake_b200 import ops as mops
from bench import c2_flops_per_unit, readme_gate_specs


def timed(fn, pool, reps=50):
    for i in range(5):
        fn(pool[i % len(pool)])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(pool[i % len(pool)])
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    n = int(os.environ.get("N", 16))
    B = 65536
    rng = np.random.default_rng(0)
    circ = mops.CompiledCircuit(n, readme_gate_specs(n, mops), 2)
    pool = [torch.as_tensor(rng.uniform(0, 2 * np.pi, (B, 2))).cuda() for _ in range(64)]
    peak = mops.probe_fp64_tflops(0, 2000)
    flops = c2_flops_per_unit(n) * B
    ref = None
    for route, env in [(1, {}), (2, {})] + [(2, {"MCB200_COV_IB": str(ib), "MCB200_COV_BPS": str(bps)})
                                            for bps in (5, 4, 3) for ib in (1, 2, 4, 8, 16, 32)]:
        for k in ("MCB200_COV_IB", "MCB200_COV_BPS"):
            os.environ.pop(k, None)
        os.environ.update(env)
        _lib.set_expval_route(route)
        ms = timed(lambda p: mops.circuit_expval_projector(circ, p), pool)
        out = mops.circuit_expval_projector(circ, pool[0]).cpu().numpy()
        if ref is None:
            ref = out
        err = float(np.max(np.abs(out - ref) / np.maximum(np.abs(ref), 1e-3)))
        print(f"route {route} {env}: {ms * 1e3:8.1f} us  {flops / ms / 1e9:6.2f} TFLOP/s (algorithmic)  frac {flops / ms / 1e9 / peak:.3f}"
              f"  max rel diff vs route 1 {err:.2e}", flush=True)
    _lib.set_expval_route(0)


if __name__ == "__main__":
    main()
