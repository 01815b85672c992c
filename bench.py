#!/usr/bin/env python
"""Benchmark of the matchgate-circuit contraction hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c4|c5] [--impl ours|reference]

One JSON line on stdout (rank 0).  A "step" is one pass of the hot path over one batch of synthetic input:
  c2 (default)  N=16 README circuit (CompRxRx/RyRy/RzRz + fSWAP), projector expval, batch 65536 per GPU
                -> circuit expectation values / s            (BASELINE.json configs[1])
  c3            N=32 FermionicPQCKernel Gram 8192 x 8192 (33 550 336 evaluated cells), rows sharded over the GPUs,
                one all-gather -> Gram entries / s           (configs[2], strong scaling)
  c4            N=64 depth-64 brickwork, qml.probs over 8 wires (256 outcomes), batch 16384 per GPU
                -> marginal distributions / s                (configs[3])
  c5            N=128 Nystroem Gram: 4096 landmarks (square) + 262144 x 4096 (transform), block-factorised
                -> Gram entries / s                          (configs[4], strong scaling)
`value` is timed with CUDA events on inputs resident in HBM (a rotating pool larger than L2 for c2); `e2e` goes through
the public API (NonInteractingFermionicDevice / FermionicPQCKernel) from pinned HOST buffers with the result read back.
`--impl reference` times the reference's CPU op sequence (oracle/ref_torch.py; the reference itself cannot be
installed in this image) on the host cores with the same metric / unit / config keys.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

L2_BYTES = 126 * 1024 * 1024
WORKLOADS = {
    "c2": "c2: N=16 README CompRxRx/RyRy/RzRz+fSWAP circuit, projector |0..0> expval, batch 65536 per GPU (BASELINE configs[1])",
    "c3": "c3: N=32 FermionicPQCKernel(rotations='X,Z') Gram 8192x8192, 33550336 evaluated cells (BASELINE configs[2])",
    "c5": "c5: N=128 FermionicPQCKernel Nystroem Gram: 4096 landmarks (K_mm) + 262144 samples x 4096 landmarks (transform), 1073737728 evaluated cells (BASELINE configs[4])",
    "c4": "c4: N=64 depth-64 brickwork matchgate circuit, probs over 8 wires (256 outcomes), batch 16384 per GPU (BASELINE configs[3])",
}
# DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture (profiles/r01/ncu_*.txt)
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "traffic.json")


# ----------------------------------------------------------------------------------------------- helpers
class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md recipe)."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.samples = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.perf_counter(), line.strip()))

    def stop(self, t0: float, t1: float):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        rows = [s for t, s in self.samples if t0 - 0.05 <= t <= t1 + 0.05] or [s for _, s in self.samples]
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            p = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(p[0]))
                mx.append(float(p[1]))
                for name, v in zip(names, p[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def readme_gate_specs(n, mops):
    """README.md:48-66: per even pair CompRxRx, CompRyRy, CompRzRz with the SAME (theta, phi); fSWAP on odd pairs."""
    g = []
    for w in range(0, n - 1, 2):
        for kind in (mops.GATE_RXRX, mops.GATE_RYRY, mops.GATE_RZRZ):
            g.append(mops.GateSpec(kind, w, w + 1, 0))
    for w in range(1, n - 1, 2):
        g.append(mops.GateSpec(mops.GATE_FSWAP, w, w + 1))
    return g


def c2_flops_per_unit(n):
    """Algorithmic FP64 flop per expectation value:
    structured chain: 3*(n/2) rotation gates, 2 Givens pairs x 6 flop per column, d columns;
    conjugation with a basis-state Lambda_0: d^3; Parlett-Reid Pfaffian exploiting skew symmetry: d^3 / 3."""
    d = 2 * n
    return 3 * (n // 2) * 12 * d + d ** 3 + d ** 3 / 3.0


def gram_flops_per_entry(d):
    """d^2 adds (Lambda_i + Lambda_j) + d^3/3 Pfaffian."""
    return d * d + d ** 3 / 3.0


# ----------------------------------------------------------------------------------------------- reference arm
def run_reference(args):
    import torch

    from oracle import kernels as okern
    from oracle import ref_torch

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    cores = torch.get_num_threads()
    rng = np.random.default_rng(0)
    if args.workload == "c2":
        n, bs = 16, 8192
        params = rng.uniform(0, 2 * np.pi, (bs, 2))
        step = lambda: ref_torch.readme_expvals(params, n)
        units, metric, unit = bs, "circuit_expvals_per_sec", "expvals/s"
        cfg = {"workload": WORKLOADS["c2"], "n_qubits": n, "batch_per_step": bs}
        sample = f"{bs} circuit instances per step (full config: 65536)"
    elif args.workload == "c3":
        n, ns = 32, 96
        x = rng.random((ns, 2 * n))
        k = okern.FermionicPQCKernelOracle(n_qubits=n, rotations="X,Z").fit(x)
        s = k.x_to_sptm(x)
        idx = np.stack(np.tril_indices(ns, -1), axis=1)
        step = lambda: ref_torch.gram_pairs(s, s, idx)
        units, metric, unit = len(idx), "gram_entries_per_sec", "entries/s"
        cfg = {"workload": WORKLOADS["c3"], "n_qubits": n, "pairs_per_step": int(len(idx))}
        sample = f"{len(idx)} Gram cells per step (full config: 33550336)"
    else:
        print(json.dumps({"impl": "reference", "unavailable": f"no CPU arm for workload {args.workload}"}))
        return
    for _ in range(max(1, min(args.warmup, 2))):
        step()
    steps = max(1, min(args.steps, 20))
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    value = units * steps / dt
    print(json.dumps({
        "impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference cannot be installed here (PennyLane/torch_pfaffian absent); this is the torch-CPU "
                "restatement of its op sequence (oracle/ref_torch.py), an optimistic stand-in"}))


# ----------------------------------------------------------------------------------------------- our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c4", "c5"])
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    defaults = {"c2": (200, 10), "c3": (3, 3), "c4": (20, 3), "c5": (3, 3)}[args.workload]
    args.steps = args.steps if args.steps is not None else defaults[0]
    args.warmup = max(3, args.warmup if args.warmup is not None else defaults[1])
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    import matchcake_b200 as mc
    from matchcake_b200 import _lib
    from matchcake_b200 import distributed as dm
    from matchcake_b200 import ops as mops

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: matchcake_b200 has no CPU path")
    rank = int(os.environ.get("RANK", "0"))
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if ws > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    rng = np.random.default_rng(0 + rank)
    K, W = args.steps, args.warmup

    def barrier():
        if ws > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if ws == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ------------------------------------------------------------------ workload definition
    if args.workload == "c2":
        n, B = 16, 65536
        circ = mops.CompiledCircuit(n, readme_gate_specs(n, mops), 2, device=dev)
        pool_n = max(8, (2 * L2_BYTES) // (B * 2 * 8) + 1)  # rotating inputs: pool > 2 x L2
        pool = [torch.as_tensor(rng.uniform(0, 2 * np.pi, (B, 2))).to(dev) for _ in range(pool_n)]
        gathered = torch.empty((ws * B,), dtype=torch.float64, device=dev) if ws > 1 else None

        def step(i):
            out = mops.circuit_expval_projector(circ, pool[i % len(pool)])
            if ws > 1:
                dist.all_gather_into_tensor(gathered, out)
            return out

        units_per_step = B * ws
        metric, unit, scaling = "circuit_expvals_per_sec", "expvals/s", "weak"
        cfg = {"workload": WORKLOADS["c2"], "n_qubits": n, "batch_per_gpu": B, "parallelism": f"batch sharded x{ws}, one all-gather of the outputs",
               "l2": f"rotating pool of {len(pool)} input buffers ({len(pool) * B * 16 / 2**20:.0f} MiB > 126 MiB L2)"}
        kernel_name = "expval_warp_kernel<4>"
        alg_flops_per_launch = c2_flops_per_unit(n) * B
        launches_per_step = 1
        # end-to-end through the device API with pinned host params and a host read of the result
        host_params = [torch.as_tensor(rng.uniform(0, 2 * np.pi, (B, 2))).pin_memory() for _ in range(4)]
        qdev = mc.NonInteractingFermionicDevice(wires=n, device=dev, output_device="cpu")
        from matchcake_b200.operations import CompRxRx, CompRyRy, CompRzRz, fSWAP
        from matchcake_b200.observables import Projector

        def circuit_ops(p):
            for w in range(0, n - 1, 2):
                yield CompRxRx(p, wires=[w, w + 1])
                yield CompRyRy(p, wires=[w, w + 1])
                yield CompRzRz(p, wires=[w, w + 1])
            for w in range(1, n - 1, 2):
                yield fSWAP(wires=[w, w + 1])

        proj = Projector(np.zeros(n, dtype=int), wires=range(n))

        def e2e_step(i):
            return qdev.execute_generator(circuit_ops(host_params[i % 4]), observable=proj, output_type="expval",
                                          reset=True)

        h2d, d2h = B * 2 * 8, B * 8
        e2e_units = B * ws
    elif args.workload == "c3":
        from matchcake_b200.ml.kernels import FermionicPQCKernel

        n, ns = 32, 8192
        x = np.random.default_rng(0).random((ns, 2 * n))  # same data on every rank: the Gram is sharded by rows
        kern = FermionicPQCKernel(n_qubits=n, rotations="X,Z", r_dtype=torch.float64, c_dtype=torch.complex128)
        kern.float32_gram = False
        kern.fit(x)
        x_dev = torch.as_tensor(x).to(dev)
        x_host = torch.as_tensor(x).pin_memory()

        def step(i):
            return kern(x_dev)

        out_host = torch.empty((ns, ns), dtype=torch.float64).pin_memory()

        def e2e_step(i):
            out_host.copy_(kern(x_host))  # host features in, Gram read back into (pinned) host memory
            return out_host

        entries = ns * (ns - 1) // 2
        units_per_step = entries
        e2e_units = entries
        metric, unit, scaling = "gram_entries_per_sec", "entries/s", "strong"
        cfg = {"workload": WORKLOADS["c3"], "n_qubits": n, "n_samples": ns, "parallelism": f"Gram rows sharded x{ws}, one all-gather",
               "l2": "per-sample covariances 268 MB > 126 MiB L2"}
        kernel_name = "gram_tiles_kernel<8,1,3,4,1>"
        alg_flops_per_launch = gram_flops_per_entry(2 * n) * entries / ws
        launches_per_step = None
        h2d, d2h = ns * 2 * n * 8, ns * ns * 8
    elif args.workload == "c5":
        from matchcake_b200.ml.kernels import FermionicPQCKernel

        n, ns, nl = 128, 262144, 4096
        x = np.random.default_rng(0).random((ns, n))  # depth_ = 1: 64 independent wire pairs (block-diagonal SPTM)
        kern = FermionicPQCKernel(n_qubits=n, rotations="X,Z", r_dtype=torch.float64, c_dtype=torch.complex128)
        kern.float32_gram = False
        kern.fit(x[:8])
        from sklearn.utils import check_random_state

        basis = x[check_random_state(0).permutation(ns)[:nl]]  # nystroem.py:88-91
        x_dev, basis_dev = torch.as_tensor(x).to(dev), torch.as_tensor(basis).to(dev)

        def step(i):
            kmm = kern(basis_dev)               # Nystroem.fit: square landmark Gram
            return kern(x_dev, basis_dev)       # Nystroem.transform: rectangular Gram (reference triangle rules)

        e2e_step = None
        entries = nl * (nl - 1) // 2 + ns * nl - nl * (nl + 1) // 2
        units_per_step = entries
        e2e_units = 0
        metric, unit, scaling = "gram_entries_per_sec", "entries/s", "strong"
        cfg = {"workload": WORKLOADS["c5"], "n_qubits": n, "n_samples": ns, "n_landmarks": nl,
               "parallelism": f"Gram rows sharded x{ws}, one all-gather",
               "algorithm": "depth_=1 ansatz => 64 independent wire pairs: K_ij = prod_c 2^-2 |Pf_4| (gram_blocks4_kernel)",
               "l2": "per-sample block data 805 MB + Gram 8.6 GB > 126 MiB L2"}
        kernel_name = "gram_blocks4_kernel"
        alg_flops_per_launch = 64 * 12.0 * (ns * nl - nl * (nl + 1) // 2) / ws
        launches_per_step = None
        h2d, d2h = 0, 0
    else:  # c4
        n, B, depth, k = 64, 16384, 64, 8
        gates = []
        off = 0
        kinds = (mops.GATE_RXRX, mops.GATE_RYRY, mops.GATE_RZRZ)
        for l in range(depth):
            for w in range(l % 2, n - 1, 2):
                gates.append(mops.GateSpec(kinds[l % 3], w, w + 1, off))
                off += 2
        circ = mops.CompiledCircuit(n, gates, off, device=dev)
        params = torch.rand((B, off), dtype=torch.float64, device=dev) * (2 * np.pi)  # 16384 x 4032 x 8 B = 528 MB > L2
        maj = torch.arange(2 * k, dtype=torch.int32, device=dev)
        states = torch.as_tensor(mc.NonInteractingFermionicDevice.states_to_binary(np.arange(2 ** k), k).astype(np.int8)).to(dev)

        def step(i):
            xcols = mops.circuit_columns(circ, params, maj)
            cov = mops.cov_basis(xcols, n)
            return mops.probs_all(cov, normalize=True)

        # end to end through the device API: one op object per gate (2016), parameters from pinned host memory
        from matchcake_b200.operations import SptmCompRxRx, SptmCompRyRy, SptmCompRzRz
        op_cls = (SptmCompRxRx, SptmCompRyRy, SptmCompRzRz)
        host_params = params.cpu().pin_memory()
        qdev = mc.NonInteractingFermionicDevice(wires=n, output_device="cpu")
        layout = [(op_cls[l % 3], w) for l in range(depth) for w in range(l % 2, n - 1, 2)]

        def e2e_step(i):
            p = host_params.to(dev, non_blocking=True).view(B, len(layout), 2)
            ops_ = [cls(p[:, j], wires=[w, w + 1]) for j, (cls, w) in enumerate(layout)]
            qdev.execute_generator(ops_, reset=True)
            return qdev.analytic_probability(list(range(k)))

        units_per_step = B * ws
        e2e_units = B * ws
        metric, unit, scaling = "marginal_distributions_per_sec", "distributions/s", "weak"
        cfg = {"workload": WORKLOADS["c4"], "n_qubits": n, "batch_per_gpu": B, "l2": "parameter tensor 528 MB > 126 MiB L2"}
        kernel_name = "circuit_columns_kernel"
        alg_flops_per_launch = len(gates) * 12 * (2 * k) * B
        launches_per_step = None
        h2d, d2h = B * off * 8, B * (2 ** k) * 8

    # ------------------------------------------------------------------ device-resident timing
    for i in range(W):
        step(i)
    barrier()
    _lib.reset_launch_count()
    sampler = ClockSampler(local_rank).start() if rank == 0 else None
    time.sleep(0.1 if rank == 0 else 0.0)
    barrier()
    t_wall0 = time.perf_counter()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    ev[0].record()
    for i in range(K):
        step(W + i)
        ev[i + 1].record()
    torch.cuda.synchronize()
    t_wall1 = time.perf_counter()
    barrier()
    launches = _lib.launch_count()
    total_ms = max_over_ranks(ev[0].elapsed_time(ev[K]))
    step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(K)]
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    value = units_per_step * K / (total_ms * 1e-3)

    # ------------------------------------------------------------------ end-to-end (host buffers)
    e2e = None
    if e2e_step is not None:
        for i in range(3):
            e2e_step(i)
        barrier()
        t0 = time.perf_counter()
        Ke = max(3, min(K, 50))
        for i in range(Ke):
            e2e_step(i)
        torch.cuda.synchronize()
        e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
        e2e = {"value": e2e_units * Ke / (e2e_ms * 1e-3), "unit": unit, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "steps": Ke, "ms_per_step": e2e_ms / Ke,
               "api": "NonInteractingFermionicDevice.execute_generator / FermionicPQCKernel.__call__ with pinned host inputs"}

    # ------------------------------------------------------------------ roofline of the dominant kernel
    roofline = None
    if rank == 0:
        dfma = mops.probe_fp64_tflops(0, 2000)
        dmma = mops.probe_fp64_tflops(1, 2000)
        if launches_per_step == 1:
            kernel_ms = float(np.mean(step_ms))  # one kernel per step: the step events bracket exactly that launch
        else:
            # time the dominant kernel alone with events around a single launch of it
            kernel_ms = time_dominant_kernel(args.workload, locals())
        achieved = alg_flops_per_launch / (kernel_ms * 1e-3) / 1e12
        traffic = None
        try:
            traffic = json.load(open(TRAFFIC_FILE)).get(args.workload if ws == 1 else "", {}).get("dram_bytes_per_launch")
        except (OSError, ValueError):
            pass
        roofline = {"kernel": kernel_name, "bound": "fp64", "achieved": achieved, "peak": dfma, "unit": "TFLOP/s",
                    "frac": achieved / dfma, "traffic": traffic,
                    "peak_source": "in-run register-resident DFMA probe (MEASURED_PEAKS.json has no FP64 entry); "
                                   f"DMMA m8n8k4 probe = {dmma:.1f} TFLOP/s",
                    "avg_launch_ms": kernel_ms, "algorithmic_flops_per_launch": alg_flops_per_launch}

    # ------------------------------------------------------------------ CPU baseline (rank 0, N = 1)
    cpu_baseline = None
    if rank == 0 and ws == 1 and not args.no_cpu_baseline and args.workload in ("c2", "c3"):
        from oracle import kernels as okern
        from oracle import ref_torch

        torch.set_num_threads(os.cpu_count() or 1)
        r2 = np.random.default_rng(1)
        if args.workload == "c2":
            bs = 16384
            p = r2.uniform(0, 2 * np.pi, (bs, 2))
            ref_torch.readme_expvals(p[:512], 16)
            t0 = time.perf_counter()
            reps = 0
            while time.perf_counter() - t0 < 10.0:
                ref_torch.readme_expvals(p, 16)
                reps += 1
            cpu_val = bs * reps / (time.perf_counter() - t0)
            sample = f"{reps} x {bs} instances of the same circuit (~10 s)"
        else:
            nsub = 128
            xs = r2.random((nsub, 64))
            ko = okern.FermionicPQCKernelOracle(n_qubits=32, rotations="X,Z").fit(xs)
            s = ko.x_to_sptm(xs)
            idx = np.stack(np.tril_indices(nsub, -1), axis=1)
            ref_torch.gram_pairs(s, s, idx[:256])
            t0 = time.perf_counter()
            reps = 0
            while time.perf_counter() - t0 < 10.0:
                ref_torch.gram_pairs(s, s, idx)
                reps += 1
            cpu_val = len(idx) * reps / (time.perf_counter() - t0)
            sample = f"{reps} x {len(idx)} Gram cells of a 128-sample sub-Gram (~10 s)"
        cpu_baseline = {"value": cpu_val, "unit": unit, "cores": torch.get_num_threads(), "kind": "port",
                        "sample": sample}

    if rank == 0:
        line = {"metric": metric, "value": value, "unit": unit, "n_gpus": ws, "steps": K, "warmup": W,
                "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": cfg, "e2e": e2e, "gpu_launches": int(launches),
                "roofline": roofline, "cpu_baseline": cpu_baseline, "clocks": clocks}
        print(json.dumps(line))
    if ws > 1:
        dist.destroy_process_group()


def time_dominant_kernel(workload, env):
    """CUDA-event time of ONE launch of the dominant kernel for multi-kernel steps (averaged over a few launches)."""
    import torch

    from matchcake_b200 import ops as mops

    if workload == "c5":
        kern, x_dev, basis_dev, ws = env["kern"], env["x_dev"], env["basis_dev"], env["ws"]
        from matchcake_b200 import distributed as dm

        b0, b1 = kern._x_to_blocks(x_dev), kern._x_to_blocks(basis_dev)
        s, e = dm.balanced_row_ranges(b0.shape[0], b1.shape[0], ws)[0]
        out = torch.zeros((b0.shape[0], b1.shape[0]), dtype=torch.float64, device=b0.device)
        fn = lambda: mops.gram_blocks4(b0, b1, mode=mops.GRAM_TRIANGLE if ws > 1 else mops.GRAM_REFERENCE,
                                       row_range=(s, e), out=out)
    elif workload == "c3":
        kern, x_dev, ws = env["kern"], env["x_dev"], env["ws"]
        # NOTE: this function runs on rank 0 only -- nothing in here may be a collective (kern._x_to_cov all-gathers
        # the per-sample covariances when several ranks are present, so the covariances are formed directly)
        cov = mops.cov_basis(kern._x_to_sptm(x_dev), kern.n_qubits)
        from matchcake_b200 import distributed as dm

        n = cov.shape[0]
        out = torch.zeros((n, n), dtype=torch.float64, device=cov.device)
        if ws > 1:  # rank 0's share of the triangle: its top and bottom row blocks (two launches)
            top, bottom = dm.folded_row_blocks(n, ws)[0]

            def fn():
                for s, e in (top, bottom):
                    if e > s:
                        mops.gram(cov, cov, 2.0 ** -32, mode=mops.GRAM_TRIANGLE, row_range=(s, e), out=out)
        else:
            fn = lambda: mops.gram(cov, cov, 2.0 ** -32, mode=mops.GRAM_REFERENCE, out=out)
    else:
        circ, params, maj = env["circ"], env["params"], env["maj"]
        fn = lambda: mops.circuit_columns(circ, params, maj)
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 2
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


if __name__ == "__main__":
    main()
