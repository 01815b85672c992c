"""Thin torch-facing layer over the C ABI (include/mcb200.h).

Every function takes CUDA ``float64`` tensors, launches hand-written sm_100a kernels from libmcb200.so on the
current torch stream and returns new tensors.  PyTorch is used only for device memory and streams.  Nothing here
computes on the CPU: a CPU tensor, a missing library or a missing GPU raises.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from . import autograd as _ag
from . import torch_ops as _t
from ._lib import CircuitStruct, Mcb200Error, check
from .torch_ops import _Floor  # noqa: F401  (re-exported: scripts / tests use ops._Floor)

class _Ops:
    """``_op.name`` = ``torch.ops.mcb200.name.default``: the resolved overload (skips the packet's per-call overload
    resolution); every launch below goes through a registered torch custom operator (torch_ops.py)."""

    def __getattr__(self, name):
        fn = getattr(torch.ops.mcb200, name).default
        setattr(self, name, fn)
        return fn


_op = _Ops()

GATE_RXRX, GATE_RYRY, GATE_RZRZ, GATE_FSWAP, GATE_BLOCK4_CONST, GATE_BLOCK4_PARAM = range(6)
FLAG_TRANSPOSE = 1
PF_ABS, PF_SIGNED = 0, 1
GRAM_REFERENCE, GRAM_FULL, GRAM_TRIANGLE = 0, 1, 2


_stream = _t._stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _f64(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise Mcb200Error(f"{name} must live on a CUDA device: matchcake_b200 has no CPU path")
    if t.dtype != torch.float64:
        t = t.to(torch.float64)
    return t.contiguous()


# ------------------------------------------------------------------------------------------------ circuits
@dataclass
class GateSpec:
    """Host-side gate record (include/mcb200.h: mcb200_gate_t)."""

    kind: int
    wire0: int
    wire1: int
    off: int = 0
    flags: int = 0

    @property
    def rows(self) -> Tuple[int, int]:
        return 2 * self.wire0, 2 * self.wire1 + 2


def schedule_layers(gates: Sequence[GateSpec]) -> Tuple[List[GateSpec], List[int]]:
    """Greedy layering that preserves the product: consecutive gates on disjoint wires share a layer."""
    layers: List[List[GateSpec]] = []
    used: set = set()
    for g in gates:
        wires = set(range(g.wire0, g.wire1 + 1))
        if not layers or (used & wires):
            layers.append([])
            used = set()
        layers[-1].append(g)
        used |= wires
    flat: List[GateSpec] = []
    ptr = [0]
    for layer in layers:
        flat.extend(layer)
        ptr.append(len(flat))
    return flat, ptr


class CompiledCircuit:
    """A gate-list circuit uploaded once to the device (gates in application order, R = G_0 G_1 ... G_{n-1})."""

    def __init__(self, n_qubits: int, gates: Sequence[GateSpec], n_param_cols: int,
                 scale: Optional[np.ndarray] = None, bias: Optional[np.ndarray] = None,
                 consts: Optional[np.ndarray] = None, device: Optional[torch.device] = None):
        if n_qubits < 1:
            raise ValueError("n_qubits must be >= 1")
        self.n_qubits = int(n_qubits)
        self.n_param_cols = int(n_param_cols)
        for g in gates:
            if not (0 <= g.wire0 < g.wire1 < n_qubits):
                raise ValueError(f"gate wires ({g.wire0},{g.wire1}) out of range for {n_qubits} qubits")
            if g.kind != GATE_FSWAP and g.wire1 != g.wire0 + 1:
                raise ValueError("two-qubit blocks act on consecutive wires")
            width = 16 if g.kind == GATE_BLOCK4_PARAM else (2 if g.kind in (GATE_RXRX, GATE_RYRY, GATE_RZRZ) else 0)
            if width and g.off + width > n_param_cols:
                raise ValueError("gate parameter offset exceeds the parameter columns")
        flat, ptr = schedule_layers(gates)
        self.gates = flat
        self.layer_ptr_host = ptr
        self.device = torch.device(device if device is not None else "cuda")
        arr = np.array([[g.kind, g.wire0, g.wire1, g.off, g.flags, 0] for g in flat], dtype=np.int32).reshape(-1, 6)
        self.gates_dev = torch.from_numpy(arr).to(self.device)
        self.layer_ptr_dev = torch.tensor(ptr, dtype=torch.int32, device=self.device)
        f = lambda a: None if a is None else torch.as_tensor(np.asarray(a, dtype=np.float64)).to(self.device)
        self._host_arrays = (scale, bias, consts)      # kept for the light-cone sub-circuits
        self._cones: dict = {}
        self.scale_dev, self.bias_dev, self.consts_dev = f(scale), f(bias), f(consts)
        if (self.scale_dev is None) != (self.bias_dev is None):
            raise ValueError("scale and bias must be given together")
        self._struct = CircuitStruct(
            gates=_ptr(self.gates_dev) if len(flat) else None, layer_ptr=_ptr(self.layer_ptr_dev),
            n_gates=len(flat), n_layers=len(ptr) - 1, n_qubits=self.n_qubits, n_param_cols=self.n_param_cols,
            scale=_ptr(self.scale_dev), bias=_ptr(self.bias_dev), consts=_ptr(self.consts_dev),
            nearest_neighbour=int(all(g.wire1 == g.wire0 + 1 for g in flat)), reserved=0)
        #: integer the torch.ops.mcb200.* circuit operators take (weak registry: this object owns the descriptor)
        self.handle = _t.register_circuit(self)

    @property
    def d(self) -> int:
        return 2 * self.n_qubits

    def validate(self) -> None:
        """Library-side check of the uploaded descriptor (mcb200_circuit_validate); raises Mcb200Error."""
        check(_lib.load().mcb200_circuit_validate(C.byref(self._struct)))

    # ------------------------------------------------------------------ backward light cone of a column panel
    def light_cone_gates(self, cols: Sequence[int]) -> List[GateSpec]:
        """Gates that can influence the columns ``cols`` of the SPTM.  X = G_0 (G_1 (... (G_{n-1} I[:, cols]))) is built
        from the LAST gate backwards; a gate whose wires carry only structurally zero rows of the current panel acts on
        zeros and is dropped (a long-range fSWAP touches every wire of its span).  Deep brickwork circuits measured on
        a few wires lose a large part of their gates this way (config C4: 768 of 2016); the result is unchanged."""
        support = np.zeros(self.n_qubits, dtype=bool)
        support[np.asarray(list(cols), dtype=np.int64) // 2] = True
        keep = []
        for g in reversed(self.gates):
            if support[g.wire0:g.wire1 + 1].any():
                support[g.wire0:g.wire1 + 1] = True
                keep.append(g)
        keep.reverse()
        return keep

    def light_cone(self, cols) -> Tuple["CompiledCircuit", torch.Tensor]:
        """(sub-circuit inside the light cone of ``cols``, ``cols`` as an int32 device tensor), cached per column set.
        ``cols``: sequence / numpy array / tensor of Majorana indices.  The sub-circuit shares the parameter layout."""
        if isinstance(cols, torch.Tensor):
            key = ("t", cols.data_ptr(), cols.numel(), cols._version, str(cols.device))
            hit = self._cones.get(key)
            if hit is not None and hit[2]() is cols:
                return hit[0], hit[1]
            host = cols.detach().cpu().numpy().astype(np.int64).ravel()
        else:
            host = np.asarray(cols, dtype=np.int64).ravel()
            key = ("h", host.tobytes())
            hit = self._cones.get(key)
            if hit is not None:
                return hit[0], hit[1]
        if host.size and (host.min() < 0 or host.max() >= self.d):
            raise ValueError(f"column index out of range for {self.d} Majorana modes")
        gates = self.light_cone_gates(host)
        scale, bias, consts = self._host_arrays
        sub = self if len(gates) == len(self.gates) else CompiledCircuit(
            self.n_qubits, gates, self.n_param_cols, scale=scale, bias=bias, consts=consts, device=self.device)
        cols_dev = torch.as_tensor(host.astype(np.int32), device=self.device)
        if len(self._cones) > 32:
            self._cones.clear()
        import weakref

        self._cones[key] = (sub, cols_dev, weakref.ref(cols) if isinstance(cols, torch.Tensor) else None)
        return sub, cols_dev

    def _params(self, params: torch.Tensor) -> torch.Tensor:
        params = _f64(params, "params")
        if params.ndim != 2 or params.shape[1] != self.n_param_cols:
            raise ValueError(f"params must have shape (B, {self.n_param_cols}), got {tuple(params.shape)}")
        return params


def _i32(t: Optional[torch.Tensor], device) -> Optional[torch.Tensor]:
    return None if t is None else t.to(device=device, dtype=torch.int32).contiguous()


def _i8(t: Optional[torch.Tensor], device) -> Optional[torch.Tensor]:
    return None if t is None else t.to(device=device, dtype=torch.int8).contiguous()


def circuit_columns(circ: CompiledCircuit, params: torch.Tensor, cols=None, prune: bool = True) -> torch.Tensor:
    """X[b] = R_b[:, cols] (B, d, m); ``cols=None`` gives the full SPTM (B, d, d).  Differentiable w.r.t. ``params``
    (adjoint: mcb200_circuit_columns_backward un-applies the gates).  ``cols`` may be a sequence, a numpy array or a
    tensor; with ``prune`` only the gates inside the backward light cone of the panel are launched
    (:meth:`CompiledCircuit.light_cone`, cached per column set) -- same result, fewer gates."""
    params = circ._params(params)
    if cols is None:
        return _op.circuit_columns(params, circ.handle, None)
    if prune:
        circ, cols = circ.light_cone(cols)
    elif not isinstance(cols, torch.Tensor):
        cols = torch.as_tensor(np.asarray(cols, dtype=np.int32))
    return _op.circuit_columns(params, circ.handle, _i32(cols, params.device))


def circuit_cov_basis(circ: CompiledCircuit, params: torch.Tensor, cols, bits: Optional[torch.Tensor] = None,
                      prune: bool = True) -> torch.Tensor:
    """cov[b] = (R_b^T Lambda_0(bits) R_b)[cols, cols] (B, m, m) for a basis-state input: K1 + K2 in one launch when the
    circuit is nearest-neighbour and the panel has at most 32 columns (the (B, d, m) panel stays in shared memory),
    ``circuit_columns`` + ``cov_basis`` otherwise and whenever a parameter requires grad."""
    if _ag.needs_grad(params):
        return cov_basis(circuit_columns(circ, params, cols, prune=prune), circ.n_qubits, bits)
    params = circ._params(params)
    if prune:
        circ, cols = circ.light_cone(cols)
    elif not isinstance(cols, torch.Tensor):
        cols = torch.as_tensor(np.asarray(cols, dtype=np.int32))
    return _op.circuit_cov_basis(params, circ.handle, _i32(cols, params.device), _i8(bits, params.device))


def circuit_expval_projector(circ: CompiledCircuit, params: torch.Tensor, init_bits: Optional[torch.Tensor] = None,
                             target_bits: Optional[torch.Tensor] = None, floor2: float = 0.0) -> torch.Tensor:
    """<y|rho_b|y> on all wires, one fused kernel (B,).  ``floor2`` > 0: sqrt(p^2 + floor2), see :func:`probs`."""
    if _ag.needs_grad(params):  # differentiable route: K1 -> K2 -> K3 as separate ops, each with its adjoint kernel
        n = circ.n_qubits
        dev = params.device
        cov = cov_basis(circuit_columns(circ, params), n, init_bits)
        tgt = torch.zeros((1, n), dtype=torch.int8, device=dev) if target_bits is None else \
            target_bits.to(device=dev, dtype=torch.int8).reshape(1, n)
        return probs(cov, tgt, floor2=floor2)[:, 0]
    params = circ._params(params)
    return _op.circuit_expval_projector(params, circ.handle, _i8(init_bits, params.device),
                                        _i8(target_bits, params.device), float(floor2))


def sptm_matmul(a: torch.Tensor, b: torch.Tensor, trans_a: bool = False, trans_b: bool = False) -> torch.Tensor:
    """Batched dense chain step C = op(a) @ op(b); a/b are (n,n) or (B,n,n) (a 2-D operand is shared).  Differentiable
    (the adjoint is two more chain steps on the same DMMA kernel)."""
    a, b = _f64(a, "a"), _f64(b, "b")
    n = a.shape[-1]
    if a.shape[-2:] != (n, n) or b.shape[-2:] != (n, n):
        raise ValueError("sptm_matmul expects square matrices of equal size")
    batch = max(a.shape[0] if a.ndim == 3 else 1, b.shape[0] if b.ndim == 3 else 1)
    for t in (a, b):
        if t.ndim == 3 and t.shape[0] != batch:
            raise ValueError("batch sizes differ")
    out = _op.sptm_matmul(a, b, bool(trans_a), bool(trans_b))
    return out if (a.ndim == 3 or b.ndim == 3) else out[0]


def matchgate_sptm(u: torch.Tensor, atol: float = 1e-6) -> torch.Tensor:
    """(B, 4, 4) complex matchgates -> (B, 4, 4) real SPTM blocks; raises if the imaginary part exceeds ``atol``."""
    if not u.is_cuda:
        raise Mcb200Error("u must live on a CUDA device: matchcake_b200 has no CPU path")
    u = u.to(torch.complex128).contiguous()
    single = u.ndim == 2
    ur = torch.view_as_real(u.reshape(-1, 4, 4)).contiguous()
    out, worst = _op.matchgate_sptm(ur)
    w = float(worst.item())
    if w > atol:
        raise ValueError(f"The single-particle transition matrix must be real, but its maximum absolute imaginary "
                         f"part is {w}, which exceeds {atol}.")
    return out[0] if single else out


def product_state_cov(amplitudes: torch.Tensor, extended: bool = False) -> torch.Tensor:
    """Per-qubit amplitudes (n, 2) / (B, n, 2) complex -> Majorana covariance (…, D, D), D = 2n (+1 if extended)."""
    if not amplitudes.is_cuda:
        raise Mcb200Error("amplitudes must live on a CUDA device: matchcake_b200 has no CPU path")
    a = amplitudes.to(torch.complex128).contiguous()
    single = a.ndim == 2
    ar = torch.view_as_real(a.reshape((-1,) + a.shape[-2:])).contiguous()
    out = _op.product_state_cov(ar, bool(extended))
    return out[0] if single else out


# ------------------------------------------------------------------------------------------------ covariance
def cov_basis(x: torch.Tensor, n_qubits: int, bits: Optional[torch.Tensor] = None,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """cov[b] = X_b^T Lambda_0(bits) X_b for X (B, d, m); ``out`` (B, m, m) contiguous float64 receives the result.
    Differentiable w.r.t. ``x`` (not with ``out``)."""
    x = _f64(x, "x")
    if x.shape[1] != 2 * n_qubits:
        raise ValueError("x must have 2*n_qubits rows")
    ib = _i8(bits, x.device)
    if out is None:
        return _op.cov_basis(x, int(n_qubits), ib)
    if _ag.needs_grad(x):
        raise NotImplementedError("cov_basis(out=...) writes in place and is inference-only")
    _op.cov_basis_out(out, x, int(n_qubits), ib)
    return out


def cov_dense(r: torch.Tensor, l0: torch.Tensor, idx: Optional[torch.Tensor] = None) -> torch.Tensor:
    """cov[b] = (R~_b^T L0 R~_b)[idx, idx]; L0 is (D,D) or (B,D,D) with D = d or d+1 (extended covariance).
    Differentiable w.r.t. ``r`` only."""
    if _ag.needs_grad(l0):
        raise NotImplementedError("cov_dense is differentiable w.r.t. the SPTM only: the initial-state covariance is a constant")
    r, l0 = _f64(r, "r"), _f64(l0, "l0")
    return _op.cov_dense(r, l0, _i32(idx, r.device))


# ------------------------------------------------------------------------------------------------ Pfaffians
def _apply_floor(v: torch.Tensor, floor2: float) -> torch.Tensor:
    """Differentiable restatement of the in-kernel floor for the autograd routes."""
    return torch.sqrt(v * v + floor2) if floor2 > 0.0 else v


def pfaffian(a: torch.Tensor, signed: bool = False, scale: float = 1.0, epsilon: float = 0.0) -> torch.Tensor:
    """Batched Pfaffian of real skew-symmetric matrices (..., m, m) -> (...).

    ``epsilon`` > 0 (magnitude mode only): ``scale * sqrt(Pf^2 + epsilon)``, the value of the reference's
    ``utils.pfaffian(sign=False, epsilon=...)`` = ``sqrt(|det A| + EPSILON)`` (utils/_pfaffian.py:110-118).
    Differentiable (adjoint 1/2 g Pf A^-T by warp Gauss-Jordan); plain Pfaffians see the skew part of their input."""
    eps = 0.0 if signed else float(epsilon)
    floor2 = float(scale) * float(scale) * eps
    grad = _ag.needs_grad(a)
    a = _f64(a, "a")
    m = a.shape[-1]
    if a.shape[-2] != m:
        raise ValueError("pfaffian expects square matrices")
    batch_shape = a.shape[:-2]
    a3 = a.reshape((int(np.prod(batch_shape)) if len(batch_shape) else 1, m, m))
    if grad:  # the in-kernel floor has no adjoint: apply it as a torch expression
        return _apply_floor(_op.pfaffian(a3, bool(signed), float(scale), 0.0), floor2).reshape(batch_shape)
    return _op.pfaffian(a3, bool(signed), float(scale), floor2).reshape(batch_shape)


def probs(cov: torch.Tensor, outcomes: torch.Tensor, normalize: bool = False, floor2: float = 0.0) -> torch.Tensor:
    """out[b, q] = 2^-k |Pf(cov[b] + Lambda_y(outcomes[q]))|; cov (B, 2k, 2k), outcomes (Q, k).

    ``floor2`` > 0 returns sqrt(p^2 + floor2) (before the optional normalisation): the reference's probabilities are
    ``sqrt(|det| + 1e-32)`` -- floor2 = 1e-32 for LookupTableStrategy (lookup_table_strategy.py:70-71), 4^-k 1e-32 for
    ProductStateProbabilityStrategy (product_state_strategy.py:211-212)."""
    grad = _ag.needs_grad(cov)
    cov = _f64(cov, "cov")
    outcomes = _i8(outcomes, cov.device)
    B, m, _ = cov.shape
    Q, k = outcomes.shape
    if m != 2 * k:
        raise ValueError("cov must be (B, 2k, 2k) for outcomes (Q, k)")
    if grad and m <= 128:
        p = _apply_floor(_op.probs(cov, outcomes, False, 0.0), floor2)
        return p / p.sum(dim=1, keepdim=True) if normalize else p
    if m > 128:
        # more than 64 measured wires: the fused gather kernels stop at 128 x 128; form cov + Lambda_y explicitly (one
        # outcome at a time) and use the large-matrix Pfaffian kernels
        sgn = outcomes.to(torch.float64) * 2.0 - 1.0                       # (Q, k): +1 for bit 1, -1 for bit 0
        ev = torch.arange(0, m, 2, device=cov.device)
        cols = []
        for q in range(Q):
            lam = torch.zeros((m, m), dtype=torch.float64, device=cov.device)
            lam[ev, ev + 1] = sgn[q]
            lam[ev + 1, ev] = -sgn[q]
            cols.append(pfaffian(cov + lam, signed=False, scale=2.0 ** -k, epsilon=floor2 * 4.0 ** k))
        out = torch.stack(cols, dim=1)
        return out / out.sum(dim=1, keepdim=True) if normalize else out
    return _op.probs(cov, outcomes, bool(normalize), float(floor2))


def probs_all(cov: torch.Tensor, normalize: bool = False, floor2: float = 0.0) -> torch.Tensor:
    """Every outcome of the k measured wires: out[b, q] for q in range(2^k), MSB first; cov (B, 2k, 2k)."""
    if _ag.needs_grad(cov):  # the elimination tree has no adjoint; enumerate the outcomes (states_to_binary order)
        k = cov.shape[-1] // 2
        q = torch.arange(1 << k, device=cov.device)
        outcomes = ((q[:, None] >> torch.arange(k - 1, -1, -1, device=cov.device)[None, :]) & 1).to(torch.int8)
        return probs(cov, outcomes, normalize, floor2=floor2)
    cov = _f64(cov, "cov")
    m = cov.shape[-1]
    if m != 2 * (m // 2) or m < 2:
        raise ValueError("cov must be (B, 2k, 2k)")
    return _op.probs_all(cov, bool(normalize), float(floor2))


def sample(cov: torch.Tensor, shots: int, seed: int = 0, forced_bits: Optional[torch.Tensor] = None,
           return_prob: bool = False):
    """Draw ``shots`` bit strings per instance from p(y) = 2^-k |Pf(cov + Lambda_y)|; cov (B, 2k, 2k).

    Returns int8 ``(shots, B, k)`` (and the probability of every drawn string ``(shots, B)`` with ``return_prob``).
    With ``forced_bits`` (shots, B, k) nothing is drawn: the probabilities of the given strings are returned."""
    cov = _f64(cov.detach(), "cov")
    bits, prob = _op.sample(cov, int(shots), _t.signed_seed(seed), _i8(forced_bits, cov.device), bool(return_prob))
    if forced_bits is not None:
        return prob
    return (bits, prob) if return_prob else bits


def sector_features(cov: torch.Tensor, index_sets: torch.Tensor) -> torch.Tensor:
    """feats[b, t] = Pf(cov[b][S_t, S_t]); cov (B, D, D), index_sets (T, s).  Differentiable w.r.t. ``cov``."""
    cov = _f64(cov, "cov")
    return _op.sector_features(cov, _i32(index_sets, cov.device))


def rowdot(feats: torch.Tensor, coeffs: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``sum_t feats[b,t] coeffs[t]``, added to ``out`` when it is given (functional: a new tensor is returned)."""
    feats = _f64(feats, "feats")
    coeffs = _f64(coeffs.to(feats.device), "coeffs")
    return _op.rowdot(feats, coeffs, None if out is None else _f64(out, "out"))


def pauli_sum_small(cov: torch.Tensor, term_ptr: torch.Tensor, indices: torch.Tensor, coeffs: torch.Tensor,
                    out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``sum_t coeffs[t] Pf(cov[b][S_t, S_t])`` for index sets of mixed sizes <= 6 (one launch), added to ``out`` when
    it is given (functional)."""
    if _ag.needs_grad(cov, out):
        raise NotImplementedError("pauli_sum_small is inference-only; use sector_features + rowdot under autograd")
    cov = _f64(cov, "cov")
    coeffs = _f64(coeffs.to(cov.device), "coeffs")
    if term_ptr.shape[0] != coeffs.shape[0] + 1:
        raise Mcb200Error(f"pauli_sum_small: term_ptr must have {coeffs.shape[0] + 1} entries, got {term_ptr.shape[0]}")
    return _op.pauli_sum_small(cov, _i32(term_ptr, cov.device), _i32(indices, cov.device), coeffs,
                               None if out is None else _f64(out, "out"))


# ------------------------------------------------------------------------------------------------ Gram
def _gram_differentiable(cov0, cov1, scale, mode, row_range, floor2=0.0, chunk: int = 8192):
    """Training-size Gram under autograd (kernel alignment, ml/kernels/kernel.py:200-283): the evaluated cells are
    batched Pfaffians of cov0[i] + cov1[j] through the differentiable Pfaffian op, ``chunk`` cells at a time (the
    reference streams ``gram_batch_size`` cells per batch; an (n_cells, d, d) tensor is never formed); cell selection,
    mirroring and the unit diagonal follow gram_matrix.py:127-161,191-195."""
    n0, n1 = cov0.shape[0], cov1.shape[0]
    rb, re = (0, n0) if row_range is None else row_range
    dev = cov0.device
    ii, jj = torch.meshgrid(torch.arange(rb, re, device=dev), torch.arange(n1, device=dev), indexing="ij")
    if mode == GRAM_FULL:
        keep = torch.ones_like(ii, dtype=torch.bool)
    else:
        keep = (jj > ii) if n0 < n1 else (ii > jj)
    ii, jj = ii[keep], jj[keep]
    chunk = max(1, int(chunk))
    parts = [_apply_floor(pfaffian(cov0[ii[s:s + chunk]] + cov1[jj[s:s + chunk]], signed=False, scale=scale), floor2)
             for s in range(0, ii.numel(), chunk)]
    vals = torch.cat(parts) if parts else torch.zeros((0,), dtype=torch.float64, device=dev)
    K = torch.zeros((n0, n1), dtype=torch.float64, device=dev).index_put((ii, jj), vals)
    if mode == GRAM_REFERENCE:
        q = min(n0, n1)
        mirror = (jj < n0) & (ii < n1)
        K = K.index_put((jj[mirror], ii[mirror]), vals[mirror])
        eye = torch.zeros_like(K)
        eye[torch.arange(q, device=dev), torch.arange(q, device=dev)] = 1.0
        K = K + eye
    return K


def _gram_buffer(out: Optional[torch.Tensor], n0: int, n1: int, rb: int, re: int, rows_only: bool, device):
    """Output buffer of a Gram launch.  ``rows_only``: just the rows [rb, re) -- shape (re - rb, n1), e.g. a view into
    the final matrix of a row-sharded build (nothing of size (n0, n1) is allocated or zero-filled); otherwise the full
    (n0, n1) matrix, zero-filled because modes TRIANGLE / REFERENCE leave cells untouched."""
    if out is not None:
        return out
    if rows_only:
        return torch.empty((re - rb, n1), dtype=torch.float64, device=device)
    return torch.zeros((n0, n1), dtype=torch.float64, device=device)


GRAM_GROWTH_FAST, GRAM_GROWTH_STRICT = 1e5, 1e3  # include/mcb200.h: mcb200_set_gram_pivot_growth
_GRAM_AUTO_SAMPLE = 32        # "auto": a sample of 32 x 32 cells is evaluated with both bounds
_GRAM_AUTO_TOL = 1e-11        # largest relative difference on the sample for which the fast bound is kept
_GRAM_AUTO_MIN_CELLS = 1 << 16  # below this the strict bound costs nothing worth a decision


class gram_pivot_growth:
    """``with gram_pivot_growth(1e3):`` -- growth bound of the static block pivots of the Gram launches inside the block
    (per thread; restored on exit)."""

    def __init__(self, growth: float):
        self.growth = float(growth)

    def __enter__(self):
        lib = _lib.load()
        self.prev = float(lib.mcb200_get_gram_pivot_growth())
        _lib.check(lib.mcb200_set_gram_pivot_growth(self.growth))
        return self

    def __exit__(self, *exc):
        _lib.load().mcb200_set_gram_pivot_growth(self.prev)
        return False


def _gram_growth_for(cov0: torch.Tensor, cov1: torch.Tensor, cells: int, pivoting) -> float:
    """``pivoting``: "fast" (bound 1e5: exact to 1e-14 on the covariances of the reference's kernel ansatz, up to 1e-8
    off in the worst of the ~2 % of cells that miss 1e-10 for generic Haar-like states), "strict" (1e3: every cell within 1e-11), a number, or
    "auto": strict for small problems, otherwise a 32 x 32 sample of the cells (evenly spaced rows and columns) is
    evaluated with both bounds and the fast one is kept only if the two agree to 1e-11 (one small device -> host read)."""
    if isinstance(pivoting, (int, float)):
        return float(pivoting)
    if pivoting == "fast":
        return GRAM_GROWTH_FAST
    if pivoting == "strict":
        return GRAM_GROWTH_STRICT
    if pivoting != "auto":
        raise ValueError(f"pivoting must be 'auto', 'fast', 'strict' or a growth bound, got {pivoting!r}")
    d = cov0.shape[-1]
    if d > 64 or cells < _GRAM_AUTO_MIN_CELLS:   # d > 64: scalar pivots guarded at 1e-3 of their row, no bound to choose
        return GRAM_GROWTH_STRICT
    n0, n1 = cov0.shape[0], cov1.shape[0]
    i0 = torch.linspace(0, n0 - 1, min(n0, _GRAM_AUTO_SAMPLE), device=cov0.device).round().long()
    i1 = torch.linspace(0, n1 - 1, min(n1, _GRAM_AUTO_SAMPLE), device=cov0.device).round().long().flip(0)
    s0, s1 = cov0.index_select(0, i0).contiguous(), cov1.index_select(0, i1).contiguous()
    both = []
    for growth in (GRAM_GROWTH_FAST, GRAM_GROWTH_STRICT):
        k = torch.empty((s0.shape[0], s1.shape[0]), dtype=torch.float64, device=cov0.device)
        with gram_pivot_growth(growth):
            _op.gram_out(k, s0, s1, 1.0, GRAM_FULL, 0, s0.shape[0], False, 0.0)
        both.append(k)
    fast, strict = both
    rel = ((fast - strict).abs() / strict.abs().clamp_min(1e-300)).max()
    return GRAM_GROWTH_FAST if float(rel) <= _GRAM_AUTO_TOL else GRAM_GROWTH_STRICT


def gram(cov0: torch.Tensor, cov1: torch.Tensor, scale: float, mode: int = GRAM_REFERENCE,
         row_range: Optional[Tuple[int, int]] = None, out: Optional[torch.Tensor] = None,
         rows_only: bool = False, floor2: float = 0.0, chunk: int = 8192, pivoting="auto") -> torch.Tensor:
    """K[i, j] = scale |Pf(cov0[i] + cov1[j])| with the reference's triangle/mirror/unit-diagonal rules.
    ``floor2`` > 0: evaluated entries are sqrt(K^2 + floor2) (the reference's Gram goes through LookupTableStrategy,
    i.e. sqrt(|det| + 1e-32), nif_kernel.py:210-221 -> lookup_table_strategy.py:70-71).  ``pivoting``: see
    :func:`_gram_growth_for`.  Under autograd the cells go through the differentiable Pfaffian op ``chunk`` at a time."""
    if _ag.needs_grad(cov0, cov1):
        return _gram_differentiable(cov0, cov1, scale, mode, row_range, floor2, chunk)
    cov0, cov1 = _f64(cov0, "cov0"), _f64(cov1, "cov1")
    n0, n1 = cov0.shape[0], cov1.shape[0]
    rb, re = (0, n0) if row_range is None else row_range
    if rows_only and mode == GRAM_REFERENCE:
        raise ValueError("rows_only needs mode TRIANGLE or FULL (REFERENCE mirrors into rows outside the block)")
    out = _gram_buffer(out, n0, n1, rb, re, rows_only, cov0.device)
    # the decision looks at the whole problem (n0 x n1), not at this call's rows: every rank of a row-sharded build
    # takes the same route
    growth = _gram_growth_for(cov0, cov1, n0 * n1, pivoting) if n0 and n1 else GRAM_GROWTH_STRICT
    with gram_pivot_growth(growth):
        _op.gram_out(out, cov0, cov1, float(scale), int(mode), int(rb), int(re), bool(rows_only), float(floor2))
    return out


def circuit_pair_blocks(circ: CompiledCircuit, params: torch.Tensor, pair_wire0: torch.Tensor,
                        init_bits: Optional[torch.Tensor] = None) -> torch.Tensor:
    """(B, C, 6) covariance blocks of a circuit whose wire pairs (w, w+1), w in ``pair_wire0``, never interact."""
    params = circ._params(params)
    return _op.circuit_pair_blocks(params, circ.handle, _i32(pair_wire0, params.device), _i8(init_bits, params.device))


def gram_blocks4(blocks0: torch.Tensor, blocks1: torch.Tensor, mode: int = GRAM_REFERENCE,
                 row_range: Optional[Tuple[int, int]] = None, out: Optional[torch.Tensor] = None,
                 rows_only: bool = False, floor2: float = 0.0) -> torch.Tensor:
    """Gram of block-diagonal covariances: blocks (n, C, 6) = upper triangles of the C 4x4 diagonal blocks."""
    blocks0, blocks1 = _f64(blocks0, "blocks0"), _f64(blocks1, "blocks1")
    n0, n1 = blocks0.shape[0], blocks1.shape[0]
    rb, re = (0, n0) if row_range is None else row_range
    if rows_only and mode == GRAM_REFERENCE:
        raise ValueError("rows_only needs mode TRIANGLE or FULL (REFERENCE mirrors into rows outside the block)")
    out = _gram_buffer(out, n0, n1, rb, re, rows_only, blocks0.device)
    _op.gram_blocks4_out(out, blocks0, blocks1, int(mode), int(rb), int(re), bool(rows_only), float(floor2))
    return out


def gram_symmetrize(k: torch.Tensor) -> torch.Tensor:
    _op.gram_symmetrize_(k)
    return k


def probe_fp64_tflops(kind: int, iters: int = 2000) -> float:
    v = _lib.load().mcb200_probe_fp64_tflops(kind, iters, _stream())
    if v < 0:
        check(1)
    return float(v)
