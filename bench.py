#!/usr/bin/env python
"""Benchmark of the matchgate-circuit contraction hot path (BASELINE.json metric: circuit expectation values / s and
kernel Gram entries / s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c4|c5|c5c] [--impl ours|reference] [--no-gram]

One JSON line on stdout (rank 0).  A "step" is one pass of the hot path over one batch of synthetic input:
  c2 (default)  N=16 README circuit (CompRxRx/RyRy/RzRz + fSWAP), projector expval, batch 65536 per GPU
                -> circuit expectation values / s            (BASELINE.json configs[1], weak scaling)
  c3            N=32 FermionicPQCKernel Gram 8192 x 8192 (33 550 336 evaluated cells), rows sharded over the GPUs,
                all-gather -> Gram entries / s               (configs[2], STRONG scaling)
  c4            N=64 depth-64 brickwork, qml.probs over 8 wires (256 outcomes), batch 16384 per GPU
                -> marginal distributions / s                (configs[3])
  c5            N=128 Nystroem Gram: 4096 landmarks (square) + 262144 x 4096 (transform), block-factorised (depth 1),
                output left row-sharded -> Gram entries / s   (configs[4], strong scaling)
  c5c           N=128 COUPLED circuit (256 features => depth 2): 512 landmarks x 2048 samples through the general
                256 x 256 Pfaffian path -> Gram entries / s
The default invocation measures c2 as the headline line AND c3 as the `gram` sub-record of the same JSON line (value,
ms_per_step, e2e, roofline, cpu_baseline at N=1, parity_max_rel, per-stage times at N>1), so that both halves of the
metric -- and the Gram's strong scaling -- are on the driver's clock at every N.

`value` is timed with CUDA events on inputs resident in HBM; `e2e` goes through the public API
(NonInteractingFermionicDevice / FermionicPQCKernel) from pinned HOST buffers with the result read back to the host.
Every timed workload's output is checked (outside the timed region) against the CPU oracle on a random sample of
>= 256 units: `parity_max_rel`; a mismatch beyond 1e-10 aborts the run.
`--impl reference` times the reference's CPU op sequence (oracle/ref_torch.py; the reference itself cannot be
installed in this image) on the host cores with the same metric / unit / config.
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
PARITY_RTOL = 1e-10
PARITY_ATOL = 1e-13   # parity-forbidden / vanishing outcomes: a relative bound is meaningless at 0
WORKLOADS = {
    "c2": "c2: N=16 README CompRxRx/RyRy/RzRz+fSWAP circuit, projector |0..0> expval, batch 65536 per GPU (BASELINE configs[1])",
    "c3": "c3: N=32 FermionicPQCKernel(rotations='X,Z') Gram 8192x8192, 33550336 evaluated cells (BASELINE configs[2])",
    "c5": "c5: N=128 FermionicPQCKernel Nystroem Gram: 4096 landmarks (K_mm) + 262144 samples x 4096 landmarks (transform), 1073737728 evaluated cells (BASELINE configs[4])",
    "c4": "c4: N=64 depth-64 brickwork matchgate circuit, probs over 8 wires (256 outcomes), batch 16384 per GPU (BASELINE configs[3])",
    "c5c": "c5c: N=128 FermionicPQCKernel with 256 features (depth 2, wire pairs coupled): Gram of 2048 samples x 512 landmarks through the general 256x256 Pfaffian path",
}
# DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` captures (profiles/)
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

    def window(self, t0: float, t1: float):
        rows = [s for t, s in self.samples if t0 - 0.05 <= t <= t1 + 0.05] or [s for _, s in self.samples]
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            p = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(p[0]))
                mx.append(float(p[1]))
                pw.append(float(p[2]))
                for name, v in zip(names, p[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons), "samples": len(sm)}

    def stop(self):
        if self.proc is not None:
            time.sleep(0.12)
            self.proc.terminate()


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


def config_for(workload: str, ws: int) -> dict:
    """The `config` object of the JSON line: identical for our arm and the reference arm."""
    if workload == "c2":
        return {"workload": WORKLOADS["c2"], "n_qubits": 16, "batch_per_gpu": 65536,
                "parallelism": f"batch sharded x{ws}, one all-gather of the outputs per step (asynchronous, under the next step's kernel)",
                "l2": "rotating pool of 253 input buffers (253 MiB > 126 MiB L2)"}
    if workload == "c3":
        return {"workload": WORKLOADS["c3"], "n_qubits": 32, "n_samples": 8192,
                "parallelism": f"Gram rows sharded x{ws} (folded row blocks), covariances and Gram all-gathered",
                "l2": "per-sample covariances 268 MB + Gram 537 MB > 126 MiB L2"}
    if workload == "c4":
        return {"workload": WORKLOADS["c4"], "n_qubits": 64, "batch_per_gpu": 16384,
                "parallelism": f"batch sharded x{ws}", "l2": "parameter tensor 528 MB > 126 MiB L2"}
    if workload == "c5":
        return {"workload": WORKLOADS["c5"], "n_qubits": 128, "n_samples": 262144, "n_landmarks": 4096,
                "parallelism": f"transform rows sharded x{ws}, K(X, landmarks) left row-sharded (SURVEY 8e)",
                "algorithm": "depth_=1 ansatz => 64 independent wire pairs: K_ij = prod_c 2^-2 |Pf_4| (gram_blocks4_kernel)",
                "l2": "per-sample block data 805 MB + Gram 8.6 GB > 126 MiB L2"}
    return {"workload": WORKLOADS["c5c"], "n_qubits": 128, "n_samples": 2048, "n_landmarks": 512,
            "parallelism": f"Gram rows sharded x{ws}", "l2": "covariances 1.3 GB > 126 MiB L2"}


def parity_max_rel(got: np.ndarray, ref: np.ndarray, atol: float = PARITY_ATOL) -> float:
    """max |got - ref| / max(|ref|, atol / rtol): the run passes when this is <= rtol, i.e. the usual
    |got - ref| <= max(rtol |ref|, atol).  Gram workloads use atol = 0 (strictly relative)."""
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    return float(np.max(np.abs(got - ref) / np.maximum(np.abs(ref), max(atol / PARITY_RTOL, 1e-300))))


# ----------------------------------------------------------------------------------------------- CPU legs
def cpu_c2(batch, seconds):
    """oracle/ref_torch.py README expvals on all host cores; returns (expvals/s, cores, sample description)."""
    import torch

    from oracle import ref_torch

    torch.set_num_threads(os.cpu_count() or 1)
    p = np.random.default_rng(1).uniform(0, 2 * np.pi, (batch, 2))
    ref_torch.readme_expvals(p[:512], 16)
    t0 = time.perf_counter()
    reps = 0
    while reps == 0 or time.perf_counter() - t0 < seconds:
        ref_torch.readme_expvals(p, 16)
        reps += 1
    dt = time.perf_counter() - t0
    return batch * reps / dt, torch.get_num_threads(), f"{reps} x {batch} instances of the c2 circuit ({dt:.1f} s)"


def cpu_c3_setup(nsub):
    from oracle import kernels as okern

    xs = np.random.default_rng(1).random((nsub, 64))
    ko = okern.FermionicPQCKernelOracle(n_qubits=32, rotations="X,Z").fit(xs)
    s = ko.x_to_sptm(xs)
    idx = np.stack(np.tril_indices(nsub, -1), axis=1)
    return s, idx


def cpu_c3(seconds, nsub=128):
    import torch

    from oracle import ref_torch

    torch.set_num_threads(os.cpu_count() or 1)
    s, idx = cpu_c3_setup(nsub)
    ref_torch.gram_pairs(s, s, idx[:256])
    t0 = time.perf_counter()
    reps = 0
    while reps == 0 or time.perf_counter() - t0 < seconds:
        ref_torch.gram_pairs(s, s, idx)
        reps += 1
    dt = time.perf_counter() - t0
    return (len(idx) * reps / dt, torch.get_num_threads(),
            f"{reps} x {len(idx)} Gram cells of a {nsub}-sample sub-Gram of the c3 kernel ({dt:.1f} s)")


def cpu_c4(seconds, batch=64):
    """numpy oracle of config C4 (dense per-layer chain, complex lookup-table read-out of the 256 outcomes)."""
    from oracle import covariance as cv
    from oracle import gates as og
    from oracle import readout

    n, depth, k = 64, 64, 8
    rng = np.random.default_rng(1)
    knames = ("CompRxRx", "CompRyRy", "CompRzRz")
    amp = cv.amplitudes_from_basis_state(np.zeros(n, int))

    def step():
        ops_ = [og.Op(knames[l % 3], (q, q + 1), params=rng.uniform(0, 2 * np.pi, (batch, 2)))
                for l in range(depth) for q in range(l % 2, n - 1, 2)]
        return readout.analytic_probability(og.chain(ops_, n), amp, list(range(k)))

    t0 = time.perf_counter()
    reps = 0
    while reps == 0 or time.perf_counter() - t0 < seconds:
        step()
        reps += 1
    dt = time.perf_counter() - t0
    return batch * reps / dt, os.cpu_count() or 1, f"{reps} x {batch} instances of the c4 circuit, numpy oracle ({dt:.1f} s)"


def cpu_c5(seconds, nf, nsub=6):
    """Per-pair reference path at N = 128 (dense R_i R_j^T, complex lookup table, 256 x 256 complex determinant): the
    reference has no block factorisation, so c5 (nf = 128) and c5c (nf = 256) cost the same per cell."""
    import torch

    from oracle import kernels as okern
    from oracle import ref_torch

    torch.set_num_threads(os.cpu_count() or 1)
    xs = np.random.default_rng(1).random((nsub, nf))
    ko = okern.FermionicPQCKernelOracle(n_qubits=128, rotations="X,Z").fit(xs)
    s = ko.x_to_sptm(xs)
    idx = np.stack(np.tril_indices(nsub, -1), axis=1)
    ref_torch.gram_pairs(s, s, idx[:4])
    t0 = time.perf_counter()
    reps = 0
    while reps == 0 or time.perf_counter() - t0 < seconds:
        ref_torch.gram_pairs(s, s, idx)
        reps += 1
    dt = time.perf_counter() - t0
    return (len(idx) * reps / dt, torch.get_num_threads(),
            f"{reps} x {len(idx)} Gram cells of a {nsub}-sample sub-Gram at N=128, {nf} features ({dt:.1f} s)")


# ----------------------------------------------------------------------------------------------- reference arm
def run_reference(args):
    """The reference's CPU op sequence (oracle/ref_torch.py) on the host cores, same metric / unit / config as ours.
    c2: every step is the FULL 65 536-instance batch of the config; gram sub-record: a bounded sample of c3's cells."""
    import torch

    from oracle import ref_torch

    rank = int(os.environ.get("RANK", "0"))
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    cores = torch.get_num_threads()
    rng = np.random.default_rng(0)
    note = ("reference cannot be installed here (PennyLane/torch_pfaffian absent); this is the torch-CPU restatement of "
            "its op sequence (oracle/ref_torch.py), an optimistic stand-in without the reference's per-gate Python")

    def timed(step, units, steps, warmup):
        for _ in range(warmup):
            step()
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        dt = time.perf_counter() - t0
        return units * steps / dt, 1e3 * dt / steps

    def gram_leg():
        s, idx = cpu_c3_setup(128)
        k = max(1, min(args.steps, 5))
        val, ms = timed(lambda: ref_torch.gram_pairs(s, s, idx), len(idx), k, 1)
        sample = f"{len(idx)} Gram cells per step (128-sample sub-Gram of c3; full config: 33550336 cells)"
        return {"metric": "gram_entries_per_sec", "value": val, "unit": "entries/s", "steps": k, "ms_per_step": ms,
                "scaling": "strong", "config": config_for("c3", ws),
                "cpu_baseline": {"value": val, "unit": "entries/s", "cores": cores, "kind": "port", "sample": sample},
                "e2e": {"value": val, "unit": "entries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}

    if args.workload == "c2":
        n, bs = 16, 65536
        params = rng.uniform(0, 2 * np.pi, (bs, 2))
        steps, warmup = max(1, min(args.steps, 20)), max(1, min(args.warmup, 2))
        val, ms = timed(lambda: ref_torch.readme_expvals(params, n), bs, steps, warmup)
        line = {"impl": "reference", "metric": "circuit_expvals_per_sec", "value": val, "unit": "expvals/s",
                "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config_for("c2", ws),
                "cpu_baseline": {"value": val, "unit": "expvals/s", "cores": cores, "kind": "port",
                                 "sample": f"{steps} steps of the full {bs}-instance batch"},
                "e2e": {"value": val, "unit": "expvals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "note": note}
        if not args.no_gram:
            line["gram"] = gram_leg()
        print(json.dumps(line))
    elif args.workload == "c3":
        g = gram_leg()
        print(json.dumps({"impl": "reference", "metric": g["metric"], "value": g["value"], "unit": g["unit"],
                          "n_gpus": args.gpus, "steps": g["steps"], "warmup": 1, "ms_per_step": g["ms_per_step"],
                          "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                          "data": "synthetic", "config": g["config"], "cpu_baseline": g["cpu_baseline"],
                          "e2e": g["e2e"], "note": note}))
    else:
        leg = {"c4": lambda: cpu_c4(10.0), "c5": lambda: cpu_c5(10.0, 128), "c5c": lambda: cpu_c5(10.0, 256)}[args.workload]
        val, used, sample = leg()
        metric, unit = (("marginal_distributions_per_sec", "distributions/s") if args.workload == "c4"
                        else ("gram_entries_per_sec", "entries/s"))
        print(json.dumps({"impl": "reference", "metric": metric, "value": val, "unit": unit, "n_gpus": args.gpus,
                          "steps": 1, "warmup": 0, "ms_per_step": None, "higher_is_better": True,
                          "scaling": "weak" if args.workload == "c4" else "strong", "vs_baseline": None, "dtype": "f64",
                          "data": "synthetic", "config": config_for(args.workload, ws),
                          "cpu_baseline": {"value": val, "unit": unit, "cores": used, "kind": "port", "sample": sample},
                          "e2e": {"value": val, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "note": note}))


# ----------------------------------------------------------------------------------------------- our arm
class Ctx:
    """Process-wide state of one bench run (rank, device, barrier / max-over-ranks helpers)."""

    def __init__(self):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.ws = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.ws > 1:
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.ws > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, ms: float) -> float:
        if self.ws == 1:
            return ms
        t = self.torch.tensor([ms], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())


def build_workload(name: str, ctx: Ctx):
    """Returns a dict: step(i), e2e_step(i) | None, units, e2e_units, metric, unit, scaling, kernel, flops_per_launch,
    launches_per_step, h2d, d2h, parity() -> float | None, dominant() -> callable launching the dominant kernel once,
    stages() -> dict | None."""
    torch = ctx.torch
    import matchcake_b200 as mc
    from matchcake_b200 import distributed as dm
    from matchcake_b200 import ops as mops

    dev, ws, rank = ctx.dev, ctx.ws, ctx.rank
    rng = np.random.default_rng(0 + rank)
    w = {"name": name, "config": config_for(name, ws), "stages": lambda: None}
    if name == "c2":
        from matchcake_b200.observables import Projector
        from matchcake_b200.operations import CompRxRx, CompRyRy, CompRzRz, fSWAP

        n, B = 16, 65536
        circ = mops.CompiledCircuit(n, readme_gate_specs(n, mops), 2, device=dev)
        pool_n = max(8, (2 * L2_BYTES) // (B * 2 * 8) + 1)  # rotating inputs: pool > 2 x L2
        pool = [torch.as_tensor(rng.uniform(0, 2 * np.pi, (B, 2))).to(dev) for _ in range(pool_n)]
        w["config"]["l2"] = f"rotating pool of {len(pool)} input buffers ({len(pool) * B * 16 / 2**20:.0f} MiB > 126 MiB L2)"
        # N > 1: a stream of batches -- the all-gather of batch i (NCCL's own stream, double-buffered output) runs under the
        # kernel of batch i + 1; `drain` (called before the final event of every timed region) waits for the last one
        gathered = [torch.empty((ws * B,), dtype=torch.float64, device=dev) for _ in range(2)] if ws > 1 else None
        last = {}
        pending = [None, None]

        def step(i):
            out = mops.circuit_expval_projector(circ, pool[i % len(pool)])
            if ws > 1:
                if pending[i & 1] is not None:
                    pending[i & 1].wait()       # the buffer about to be overwritten has been delivered
                pending[i & 1] = ctx.dist.all_gather_into_tensor(gathered[i & 1], out, async_op=True)
            last["i"], last["out"] = i % len(pool), out
            return out

        def drain():
            for k in range(2):
                if pending[k] is not None:
                    pending[k].wait()
                    pending[k] = None

        w["drain"] = drain

        host_params = [torch.as_tensor(rng.uniform(0, 2 * np.pi, (B, 2))).pin_memory() for _ in range(4)]
        qdev = mc.NonInteractingFermionicDevice(wires=n, device=dev, output_device="cpu")

        def circuit_ops(p):
            for q in range(0, n - 1, 2):
                yield CompRxRx(p, wires=[q, q + 1])
                yield CompRyRy(p, wires=[q, q + 1])
                yield CompRzRz(p, wires=[q, q + 1])
            for q in range(1, n - 1, 2):
                yield fSWAP(wires=[q, q + 1])

        proj = Projector(np.zeros(n, dtype=int), wires=range(n))

        def e2e_step(i):
            r = qdev.execute_generator(circuit_ops(host_params[i % 4]), observable=proj, output_type="expval", reset=True)
            last["e2e_i"], last["e2e_out"] = i % 4, r
            return r

        def parity():
            from oracle import gates as og
            from oracle import readout

            sel = np.random.default_rng(7).choice(B, 256, replace=False)
            zeros = np.zeros(n, dtype=int)

            def ref_of(p):
                ops_ = []
                for q in range(0, n - 1, 2):
                    for nm in ("CompRxRx", "CompRyRy", "CompRzRz"):
                        ops_.append(og.Op(nm, (q, q + 1), params=p))
                for q in range(1, n - 1, 2):
                    ops_.append(og.Op("fSWAP", (q, q + 1)))
                return readout.probs_lookup_table(og.chain(ops_, n), zeros, zeros, np.arange(n))

            # device-timed path: exact |Pf| (no floor at the ops level) -> compare with the floor removed from the oracle
            p_dev = pool[last["i"]][torch.as_tensor(sel, device=dev)].cpu().numpy()
            ref = np.sqrt(np.maximum(ref_of(p_dev) ** 2 - 1e-32, 0.0))
            rel = parity_max_rel(last["out"][torch.as_tensor(sel, device=dev)].cpu().numpy(), ref)
            # end-to-end path through the device (reference floor 1e-32 included)
            p_host = host_params[last["e2e_i"]][torch.as_tensor(sel)].numpy()
            rel = max(rel, parity_max_rel(np.asarray(last["e2e_out"])[sel], ref_of(p_host)))
            return rel

        w.update(step=step, e2e_step=e2e_step, units=B * ws, e2e_units=B * ws, metric="circuit_expvals_per_sec",
                 unit="expvals/s", scaling="weak", kernel="expval_cov_kernel<4>",
                 flops_per_launch=c2_flops_per_unit(n) * B, launches_per_step=1, h2d=B * 2 * 8, d2h=B * 8,
                 parity=parity, dominant=None, e2e_api="NonInteractingFermionicDevice.execute_generator (pinned host "
                 "parameters in, host tensor out)")
    elif name == "c3":
        from matchcake_b200.ml.kernels import FermionicPQCKernel

        n, ns = 32, 8192
        x = np.random.default_rng(0).random((ns, 2 * n))  # same data on every rank: the Gram is sharded by rows
        kern = FermionicPQCKernel(n_qubits=n, rotations="X,Z", r_dtype=torch.float64, c_dtype=torch.complex128)
        kern.float32_gram = False
        kern.fit(x)
        x_dev = torch.as_tensor(x).to(dev)
        x_host = torch.as_tensor(x).pin_memory()
        last = {}

        def step(i):
            last["out"] = kern(x_dev)
            return last["out"]

        out_host = torch.empty((ns, ns), dtype=torch.float64).pin_memory() if rank == 0 else None

        def e2e_step(i):
            k = kern(x_host)           # host features in; the assembled Gram is read back by rank 0 (pinned host memory)
            if rank == 0:
                out_host.copy_(k)
            return out_host

        entries = ns * (ns - 1) // 2

        def parity():
            from oracle import kernels as okern

            r7 = np.random.default_rng(7)
            ii, jj = r7.integers(0, ns, 300), r7.integers(0, ns, 300)
            keep = ii != jj
            ii, jj = ii[keep][:256], jj[keep][:256]
            o = okern.FermionicPQCKernelOracle(n_qubits=n, rotations="X,Z").fit(x)
            uniq, inv = np.unique(np.concatenate([ii, jj]), return_inverse=True)
            s = o.x_to_sptm(x[uniq])
            idx = np.stack([inv[: len(ii)], inv[len(ii):]], axis=1)
            ref = o.pair_expvals(s, s, idx)
            k = last["out"]
            got = k[torch.as_tensor(ii, device=dev), torch.as_tensor(jj, device=dev)].cpu().numpy()
            rel = parity_max_rel(got, ref, atol=0.0)
            # structure: symmetric, unit diagonal (gram_matrix.py:127-161)
            assert float((k - k.T).abs().max()) == 0.0 and float((torch.diagonal(k) - 1.0).abs().max()) == 0.0
            if rank == 0:
                got_h = out_host[torch.as_tensor(ii), torch.as_tensor(jj)].numpy()
                rel = max(rel, parity_max_rel(got_h, ref, atol=0.0))
            return rel

        def dominant():
            # rank-local: nothing in here may be a collective
            cov = mops.cov_basis(kern._x_to_sptm(x_dev), kern.n_qubits)
            # the kernel alone: the route the API's "auto" decision takes on this data (decided here, outside the timing)
            growth = mops._gram_growth_for(cov, cov, ns * ns, "auto")
            w["config"]["gram_pivot_growth"] = growth
            if ws > 1:
                (t0, t1), (b0, b1) = dm.folded_row_blocks(ns, ws)[rank]
                buf = torch.empty((max(t1 - t0, b1 - b0), ns), dtype=torch.float64, device=dev)

                def fn():
                    for s0, e0 in ((t0, t1), (b0, b1)):
                        if e0 > s0:
                            mops.gram(cov, cov, 2.0 ** -32, mode=mops.GRAM_TRIANGLE, row_range=(s0, e0),
                                      out=buf[: e0 - s0], rows_only=True, pivoting=growth)
                return fn
            out = torch.empty((ns, ns), dtype=torch.float64, device=dev)
            return lambda: mops.gram(cov, cov, 2.0 ** -32, mode=mops.GRAM_REFERENCE, out=out, pivoting=growth)

        def stages():
            """Per-stage device times of one sharded build (ms, this rank): names the residual next to the kernel."""
            kern.timeline = []
            e0 = dm._event()
            kern(x_dev)
            torch.cuda.synchronize()
            tl, kern.timeline = kern.timeline, None
            out, prev = {}, e0
            for nm, ev in tl:
                if nm == "begin":
                    out["precompute_other"] = out.get("precompute_other", 0.0) + prev.elapsed_time(ev)
                else:
                    out[nm] = prev.elapsed_time(ev)
                prev = ev
            return out

        w.update(step=step, e2e_step=e2e_step, units=entries, e2e_units=entries, metric="gram_entries_per_sec",
                 unit="entries/s", scaling="strong", kernel="gram_tiles_flat_kernel<8,0,2,4>",
                 flops_per_launch=gram_flops_per_entry(2 * n) * entries / ws, launches_per_step=None,
                 h2d=ns * 2 * n * 8, d2h=ns * ns * 8, parity=parity, dominant=dominant,
                 stages=stages if ws > 1 else (lambda: None),
                 e2e_api="FermionicPQCKernel.__call__ (pinned host features in, Gram copied to pinned host memory on rank 0)")
    elif name in ("c5", "c5c"):
        from sklearn.utils import check_random_state

        from matchcake_b200.ml.kernels import FermionicPQCKernel

        if name == "c5":
            n, ns, nl, nf = 128, 262144, 4096, 128
        else:
            n, ns, nl, nf = 128, 2048, 512, 256
        x = np.random.default_rng(0).random((ns, nf))  # nf = 128: depth_ = 1 (block-diagonal SPTM); 256: depth 2
        kern = FermionicPQCKernel(n_qubits=n, rotations="X,Z", r_dtype=torch.float64, c_dtype=torch.complex128)
        kern.float32_gram = False
        kern.fit(x[:8])
        land_idx = check_random_state(0).permutation(ns)[:nl]  # nystroem.py:88-91
        basis = x[land_idx]
        x_dev, basis_dev = torch.as_tensor(x).to(dev), torch.as_tensor(basis).to(dev)
        last = {}

        def step(i):
            if name == "c5":
                last["kmm"] = kern(basis_dev)                          # Nystroem.fit: square landmark Gram
            last["knm"] = kern(x_dev, basis_dev, gather=False)        # Nystroem.transform: rectangle, row-sharded
            return last["knm"]

        tri = nl * (nl - 1) // 2
        rect = ns * nl - nl * (nl + 1) // 2
        entries = (tri if name == "c5" else 0) + rect
        e2e_step = None
        if name == "c5c":
            # end to end through the kernel's public call: pinned host features in, Gram read back to pinned host memory.
            # (c5 has no e2e leg: its 262144 x 4096 float64 output is 8.6 GB per step -- Nystroem.transform keeps it
            # row-sharded on the devices and only the (n, 4096) embedding leaves them.)
            x_host, basis_host = torch.as_tensor(x).pin_memory(), torch.as_tensor(basis).pin_memory()
            out_host = torch.empty((ns, nl), dtype=torch.float64).pin_memory() if rank == 0 else None

            def e2e_step(i):
                k = kern(x_host, basis_host)
                if rank == 0:
                    out_host.copy_(k)
                return out_host

        def parity():
            from oracle import kernels as okern

            knm = last["knm"]
            loc, (s, e) = knm if isinstance(knm, tuple) else (knm, (0, ns))
            r7 = np.random.default_rng(7 + rank)
            cnt = 256 if name == "c5" else 24   # 256 x 256 complex determinants in the oracle: keep c5c short
            ii = r7.integers(max(s, nl), e, cnt) if e > max(s, nl) else np.zeros(0, dtype=int)
            jj = r7.integers(0, nl, len(ii))
            if len(ii) == 0:
                return 0.0
            o = okern.FermionicPQCKernelOracle(n_qubits=n, rotations="X,Z").fit(x[:8])
            ui, inv_i = np.unique(ii, return_inverse=True)
            uj, inv_j = np.unique(jj, return_inverse=True)
            s0, s1 = o.x_to_sptm(x[ui]), o.x_to_sptm(basis[uj])
            pairs = np.stack([inv_i, inv_j], axis=1)
            ref = o.pair_expvals(s0, s1, pairs)
            got = loc[torch.as_tensor(ii - s, device=dev), torch.as_tensor(jj, device=dev)].cpu().numpy()
            rel = parity_max_rel(got, ref, atol=0.0)
            # At N = 128 the entries (1e-17 .. 1e-20) sit below the reference's floor sqrt(K^2 + 1e-32), which both sides
            # apply: the comparison above is dominated by it.  Also compare the EXACT |Pf| (floor off, same kernels,
            # rank-local launch) of the same cells with the oracle's log-determinant.
            from oracle import lookup_table as lt

            zeros = np.zeros(n, dtype=int)
            r = np.einsum("pij,pkj->pik", s0[pairs[:, 0]], s1[pairs[:, 1]])
            obs = lt.observable_matrix(lt.transition_matrix(r), zeros, zeros, np.arange(n))
            obs = obs.reshape((-1,) + obs.shape[-2:])
            exact_ref = np.exp(0.5 * np.linalg.slogdet(obs)[1])
            xs, bs = torch.as_tensor(x[ui]).to(dev), torch.as_tensor(basis[uj]).to(dev)
            if name == "c5":
                sub = mops.gram_blocks4(kern._x_to_blocks(xs), kern._x_to_blocks(bs), mode=mops.GRAM_FULL)
            else:
                sub = mops.gram(mops.cov_basis(kern._x_to_sptm(xs), n), mops.cov_basis(kern._x_to_sptm(bs), n), 2.0 ** -n,
                                mode=mops.GRAM_FULL)
            exact = sub[torch.as_tensor(inv_i, device=dev), torch.as_tensor(inv_j, device=dev)].cpu().numpy()
            last["exact_rel"] = parity_max_rel(exact, exact_ref, atol=0.0)
            return max(rel, last["exact_rel"])

        def dominant():
            if name == "c5":
                b0, b1 = kern._x_to_blocks(x_dev), kern._x_to_blocks(basis_dev)
                s, e = dm.balanced_row_ranges(ns, nl, ws)[rank]
                out = torch.empty((e - s, nl), dtype=torch.float64, device=dev)
                return lambda: mops.gram_blocks4(b0, b1, mode=mops.GRAM_TRIANGLE, row_range=(s, e), out=out, rows_only=True)
            c0 = mops.cov_basis(kern._x_to_sptm(x_dev), n)
            c1 = mops.cov_basis(kern._x_to_sptm(basis_dev), n)
            s, e = dm.balanced_row_ranges(ns, nl, ws)[rank]
            out = torch.empty((e - s, nl), dtype=torch.float64, device=dev)
            return lambda: mops.gram(c0, c1, 2.0 ** -n, mode=mops.GRAM_TRIANGLE, row_range=(s, e), out=out, rows_only=True)

        if name == "c5":
            flops = 64 * 12.0 * rect / ws
            kernel_name = "gram_blocks4_kernel"
        else:
            flops = gram_flops_per_entry(2 * n) * rect / ws
            kernel_name = "gram_panel_kernel"
        w.update(step=step, e2e_step=e2e_step, units=entries, e2e_units=rect if e2e_step else 0,
                 metric="gram_entries_per_sec", unit="entries/s", scaling="strong", kernel=kernel_name,
                 flops_per_launch=flops, launches_per_step=None, h2d=(ns + nl) * nf * 8 if e2e_step else 0,
                 d2h=ns * nl * 8 if e2e_step else 0, parity=parity, dominant=dominant,
                 e2e_api="FermionicPQCKernel.__call__(x, landmarks) (pinned host features in, Gram copied to pinned "
                         "host memory on rank 0)" if e2e_step else None)
    else:  # c4
        from matchcake_b200.operations import SptmCompRxRx, SptmCompRyRy, SptmCompRzRz

        n, B, depth, k = 64, 16384, 64, 8
        gates = []
        off = 0
        kinds = (mops.GATE_RXRX, mops.GATE_RYRY, mops.GATE_RZRZ)
        knames = ("CompRxRx", "CompRyRy", "CompRzRz")
        names = []
        for l in range(depth):
            for q in range(l % 2, n - 1, 2):
                gates.append(mops.GateSpec(kinds[l % 3], q, q + 1, off))
                names.append((knames[l % 3], q, off))
                off += 2
        circ = mops.CompiledCircuit(n, gates, off, device=dev)
        params = torch.rand((B, off), dtype=torch.float64, device=dev) * (2 * np.pi)  # 16384 x 4032 x 8 B = 528 MB > L2
        maj = torch.arange(2 * k, dtype=torch.int32, device=dev)
        last = {}

        def step(i):
            xcols = mops.circuit_columns(circ, params, maj)
            cov = mops.cov_basis(xcols, n)
            last["out"] = mops.probs_all(cov, normalize=True)
            return last["out"]

        op_cls = (SptmCompRxRx, SptmCompRyRy, SptmCompRzRz)
        host_params = params.cpu().pin_memory()
        qdev = mc.NonInteractingFermionicDevice(wires=n, output_device="cpu")
        layout = [(op_cls[l % 3], q) for l in range(depth) for q in range(l % 2, n - 1, 2)]

        def e2e_step(i):
            # one (B, P) parameter tensor + gate layout; H2D chunks double-buffered against the kernels (device.py)
            last["e2e_out"] = qdev.execute_layered(layout, host_params, probs_wires=list(range(k)))
            return last["e2e_out"]

        def parity():
            from oracle import covariance as cv
            from oracle import gates as og
            from oracle import readout

            sel = [0, 7777, B - 1]
            pp = params[sel].cpu().numpy()
            ops_ = [og.Op(nm, (q, q + 1), params=pp[:, o:o + 2]) for nm, q, o in names]
            ref = readout.analytic_probability(og.chain(ops_, n), cv.amplitudes_from_basis_state(np.zeros(n, int)),
                                               list(range(k)))
            rel = parity_max_rel(last["out"][sel].cpu().numpy(), ref)
            if "e2e_out" in last:
                rel = max(rel, parity_max_rel(np.asarray(last["e2e_out"])[sel], ref))
            return rel

        w.update(step=step, e2e_step=e2e_step, units=B * ws, e2e_units=B * ws, metric="marginal_distributions_per_sec",
                 unit="distributions/s", scaling="weak", kernel="circuit_columns_kernel",
                 flops_per_launch=len(gates) * 12 * (2 * k) * B, launches_per_step=None, h2d=B * off * 8,
                 d2h=B * (2 ** k) * 8, parity=parity,
                 dominant=lambda: (lambda: mops.circuit_columns(circ, params, maj)),
                 e2e_api="NonInteractingFermionicDevice.execute_layered (pinned host parameters streamed in chunks, "
                 "host probabilities out)",
                 hbm_bytes_per_launch=B * off * 8 + B * 2 * n * 2 * k * 8)
    return w


def measure(w, ctx: Ctx, K: int, W: int, sampler, peaks, cpu_leg, sustained_s: float = 0.0):
    """Times one workload: device-resident K steps (CUDA events, max over ranks), end-to-end, dominant kernel / roofline,
    parity sample, optional CPU baseline and sustained window.  Returns the record dict (meaningful on rank 0)."""
    torch = ctx.torch
    from matchcake_b200 import _lib

    step = w["step"]
    for i in range(W):
        step(i)
    ctx.barrier()
    _lib.reset_launch_count()
    t_wall0 = time.perf_counter()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    drain = w.get("drain", lambda: None)
    drain()  # nothing of the warm-up is still in flight
    ev[0].record()
    for i in range(K):
        step(W + i)
        if i == K - 1:
            drain()  # asynchronous collectives of the last steps end inside the timed region
        ev[i + 1].record()
    torch.cuda.synchronize()
    t_wall1 = time.perf_counter()
    ctx.barrier()
    launches = _lib.launch_count()
    total_ms = ctx.max_over_ranks(ev[0].elapsed_time(ev[K]))
    step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(K)]
    rec = {"metric": w["metric"], "value": w["units"] * K / (total_ms * 1e-3), "unit": w["unit"], "steps": K,
           "warmup": W, "ms_per_step": total_ms / K, "scaling": w["scaling"], "config": w["config"],
           "gpu_launches": int(launches)}
    if sampler is not None:
        rec["clocks"] = sampler.window(t_wall0, t_wall1)

    # sustained window (>= sustained_s seconds of back-to-back steps): numerator of the sustained roofline pair
    sustained = None
    if sustained_s > 0:
        reps = max(K, int(sustained_s * 1e3 / max(total_ms / K, 1e-3)) + 1)
        ctx.barrier()
        tw0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(reps):
            step(i)
        drain()
        e1.record()
        torch.cuda.synchronize()
        tw1 = time.perf_counter()
        s_ms = ctx.max_over_ranks(e0.elapsed_time(e1))
        sustained = {"seconds": s_ms * 1e-3, "steps": reps, "ms_per_step": s_ms / reps,
                     "value": w["units"] * reps / (s_ms * 1e-3)}
        if sampler is not None:
            sustained["clocks"] = sampler.window(tw0, tw1)

    # end-to-end (host buffers in, host result out)
    if w["e2e_step"] is not None:
        for i in range(3):
            w["e2e_step"](i)
        ctx.barrier()
        t0 = time.perf_counter()
        Ke = max(3, min(K, 50))
        for i in range(Ke):
            w["e2e_step"](i)
        torch.cuda.synchronize()
        e2e_ms = ctx.max_over_ranks((time.perf_counter() - t0) * 1e3)
        rec["e2e"] = {"value": w["e2e_units"] * Ke / (e2e_ms * 1e-3), "unit": w["unit"], "h2d_bytes_per_step": w["h2d"],
                      "d2h_bytes_per_step": w["d2h"], "steps": Ke, "ms_per_step": e2e_ms / Ke, "api": w["e2e_api"]}
    else:
        rec["e2e"] = None

    # per-stage times of the sharded build (collective: every rank takes part)
    st = w["stages"]()
    if st:
        rec["stages_ms_rank0"] = st

    # parity of the timed outputs (rank-local sample vs the CPU oracle), max over ranks
    rel = w["parity"]()
    rel = ctx.max_over_ranks(rel)
    rec["parity_max_rel"] = rel
    rec["parity"] = {"checker": "oracle/ (numpy restatement of the reference)", "rtol": PARITY_RTOL,
                     "atol": 0.0 if w["metric"] == "gram_entries_per_sec" else PARITY_ATOL,
                     "sample": ">= 256 random units of the last timed step (and of the last e2e step)"}
    if not rel <= PARITY_RTOL:
        raise SystemExit(f"bench.py: parity check of workload {w['name']} failed: max rel {rel:.3e} > {PARITY_RTOL}")

    # roofline of the dominant kernel (rank 0's launch; every rank runs the same amount of work)
    if w["launches_per_step"] == 1:
        kernel_ms = float(np.mean(step_ms))  # one kernel per step: the step events bracket exactly that launch
    else:
        fn = w["dominant"]()
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(2):
            fn()
        e1.record()
        torch.cuda.synchronize()
        kernel_ms = e0.elapsed_time(e1) / 2
    if ctx.rank == 0:
        achieved = w["flops_per_launch"] / (kernel_ms * 1e-3) / 1e12
        traffic = None
        try:
            traffic = json.load(open(TRAFFIC_FILE)).get(w["name"] if ctx.ws == 1 else "", {}).get("dram_bytes_per_launch")
        except (OSError, ValueError):
            pass
        rec["roofline"] = {"kernel": w["kernel"], "bound": "fp64", "achieved": achieved, "peak": peaks["dfma"],
                           "unit": "TFLOP/s", "frac": achieved / peaks["dfma"], "traffic": traffic,
                           "peak_source": "in-run register-resident DFMA probe, burst (MEASURED_PEAKS.json has no FP64 "
                                          f"entry); DMMA m8n8k4 probe = {peaks['dmma']:.1f} TFLOP/s",
                           "avg_launch_ms": kernel_ms, "algorithmic_flops_per_launch": w["flops_per_launch"]}
        if "hbm_bytes_per_launch" in w:
            gbs = w["hbm_bytes_per_launch"] / (kernel_ms * 1e-3) / 1e9
            rec["roofline"]["hbm"] = {"achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                      "frac": gbs / peaks["hbm_gbs"], "algorithmic_bytes_per_launch": w["hbm_bytes_per_launch"]}
        if sustained is not None:
            s_ach = w["flops_per_launch"] / (sustained["ms_per_step"] * 1e-3) / 1e12 if w["launches_per_step"] == 1 else None
            sustained["roofline"] = None if s_ach is None else {
                "achieved": s_ach, "peak": peaks["dfma_sustained"], "unit": "TFLOP/s",
                "frac": s_ach / peaks["dfma_sustained"],
                "peak_source": "DFMA probe run back to back for >= 1 s per repetition (sustained clocks)"}
            rec["sustained"] = sustained
        if cpu_leg is not None:
            val, cores, sample = cpu_leg()
            rec["cpu_baseline"] = {"value": val, "unit": w["unit"], "cores": cores, "kind": "port", "sample": sample}
        else:
            rec["cpu_baseline"] = None
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c4", "c5", "c5c"])
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gram", action="store_true", help="skip the c3 `gram` sub-record of the default (c2) line")
    ap.add_argument("--no-sustained", action="store_true")
    args = ap.parse_args()
    defaults = {"c2": (200, 10), "c3": (3, 3), "c4": (20, 3), "c5": (3, 3), "c5c": (3, 3)}[args.workload]
    args.steps = args.steps if args.steps is not None else defaults[0]
    args.warmup = max(3, args.warmup if args.warmup is not None else defaults[1])
    if args.impl == "reference":
        return run_reference(args)

    import torch

    from matchcake_b200 import ops as mops

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: matchcake_b200 has no CPU path")
    ctx = Ctx()
    K, W = args.steps, args.warmup
    peaks = {"hbm_gbs": 6538.3}
    try:
        peaks["hbm_gbs"] = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except (OSError, ValueError, KeyError):
        pass
    if ctx.rank == 0:
        peaks["dfma"] = mops.probe_fp64_tflops(0, 2000)
        peaks["dmma"] = mops.probe_fp64_tflops(1, 2000)
        peaks["dfma_sustained"] = None if args.no_sustained else mops.probe_fp64_tflops(0, 1000000)
    sampler = ClockSampler(ctx.local_rank).start() if ctx.rank == 0 else None
    time.sleep(0.1 if ctx.rank == 0 else 0.0)

    cpu_ok = ctx.rank == 0 and ctx.ws == 1 and not args.no_cpu_baseline
    cpu_legs = {"c2": lambda: cpu_c2(65536, 10.0), "c3": lambda: cpu_c3(10.0), "c4": lambda: cpu_c4(10.0),
                "c5": lambda: cpu_c5(10.0, 128), "c5c": lambda: cpu_c5(10.0, 256)}
    primary = build_workload(args.workload, ctx)
    rec = measure(primary, ctx, K, W, sampler, peaks, cpu_legs.get(args.workload) if cpu_ok else None,
                  sustained_s=0.0 if (args.no_sustained or args.workload != "c2") else 2.0)
    del primary
    gram = None
    if args.workload == "c2" and not args.no_gram:
        torch.cuda.empty_cache()
        g = build_workload("c3", ctx)
        gram = measure(g, ctx, max(1, min(K, 5)), 3, sampler, peaks, cpu_legs["c3"] if cpu_ok else None)
        del g
    if sampler is not None:
        sampler.stop()
    if ctx.rank == 0:
        line = {"metric": rec["metric"], "value": rec["value"], "unit": rec["unit"], "n_gpus": ctx.ws, "steps": K,
                "warmup": W, "ms_per_step": rec["ms_per_step"], "higher_is_better": True, "scaling": rec["scaling"],
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": rec["config"], "e2e": rec["e2e"],
                "gpu_launches": rec["gpu_launches"] + (gram["gpu_launches"] if gram else 0),
                "roofline": rec.get("roofline"), "cpu_baseline": rec.get("cpu_baseline"), "clocks": rec.get("clocks"),
                "parity_max_rel": rec["parity_max_rel"], "parity": rec["parity"]}
        for key in ("sustained", "stages_ms_rank0"):
            if key in rec:
                line[key] = rec[key]
        if gram is not None:
            gram["higher_is_better"] = True
            gram["n_gpus"] = ctx.ws
            line["gram"] = gram
        print(json.dumps(line))
    if ctx.ws > 1:
        ctx.dist.destroy_process_group()


if __name__ == "__main__":
    main()
