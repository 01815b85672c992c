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
from ._lib import CircuitStruct, Mcb200Error, check

GATE_RXRX, GATE_RYRY, GATE_RZRZ, GATE_FSWAP, GATE_BLOCK4_CONST, GATE_BLOCK4_PARAM = range(6)
FLAG_TRANSPOSE = 1
PF_ABS, PF_SIGNED = 0, 1
GRAM_REFERENCE, GRAM_FULL, GRAM_TRIANGLE = 0, 1, 2


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


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


class _Floor:
    """``with _Floor(floor2):`` -- magnitude floor of the launches inside the block (include/mcb200.h:
    mcb200_set_abs_floor; the option is per thread and always restored to 0 = off)."""

    def __init__(self, floor2: float):
        self.floor2 = float(floor2)

    def __enter__(self):
        if self.floor2 != 0.0:
            check(_lib.load().mcb200_set_abs_floor(self.floor2))

    def __exit__(self, *exc):
        if self.floor2 != 0.0:
            _lib.load().mcb200_set_abs_floor(0.0)
        return False


def _workspace(nbytes: int, device) -> Optional[torch.Tensor]:
    if nbytes <= 0:
        return None
    return torch.empty((nbytes + 7) // 8, dtype=torch.float64, device=device)


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
        self.scale_dev, self.bias_dev, self.consts_dev = f(scale), f(bias), f(consts)
        if (self.scale_dev is None) != (self.bias_dev is None):
            raise ValueError("scale and bias must be given together")
        self._struct = CircuitStruct(
            gates=_ptr(self.gates_dev) if len(flat) else None, layer_ptr=_ptr(self.layer_ptr_dev),
            n_gates=len(flat), n_layers=len(ptr) - 1, n_qubits=self.n_qubits, n_param_cols=self.n_param_cols,
            scale=_ptr(self.scale_dev), bias=_ptr(self.bias_dev), consts=_ptr(self.consts_dev),
            nearest_neighbour=int(all(g.wire1 == g.wire0 + 1 for g in flat)), reserved=0)

    @property
    def d(self) -> int:
        return 2 * self.n_qubits

    def _params(self, params: torch.Tensor) -> torch.Tensor:
        params = _f64(params, "params")
        if params.ndim != 2 or params.shape[1] != self.n_param_cols:
            raise ValueError(f"params must have shape (B, {self.n_param_cols}), got {tuple(params.shape)}")
        return params


def circuit_columns(circ: CompiledCircuit, params: torch.Tensor, cols: Optional[torch.Tensor] = None) -> torch.Tensor:
    """X[b] = R_b[:, cols] (B, d, m); ``cols=None`` gives the full SPTM (B, d, d)."""
    if _ag.needs_grad(params):
        return _ag.CircuitColumns.apply(params, circ, cols)
    params = circ._params(params)
    B = params.shape[0]
    if cols is not None:
        cols = cols.to(device=params.device, dtype=torch.int32).contiguous()
        m = cols.numel()
    else:
        m = circ.d
    out = torch.empty((B, circ.d, m), dtype=torch.float64, device=params.device)
    check(_lib.load().mcb200_circuit_columns(C.byref(circ._struct), _ptr(params), B, _ptr(cols), m, _ptr(out),
                                             _stream()))
    return out


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
    B = params.shape[0]
    bits = lambda t: None if t is None else t.to(device=params.device, dtype=torch.int8).contiguous()
    ib, tb = bits(init_bits), bits(target_bits)
    out = torch.empty((B,), dtype=torch.float64, device=params.device)
    with _Floor(floor2):
        check(_lib.load().mcb200_circuit_expval_projector(C.byref(circ._struct), _ptr(params), B, _ptr(ib), _ptr(tb),
                                                          _ptr(out), _stream()))
    return out


def sptm_matmul(a: torch.Tensor, b: torch.Tensor, trans_a: bool = False, trans_b: bool = False) -> torch.Tensor:
    """Batched dense chain step C = op(a) @ op(b); a/b are (n,n) or (B,n,n) (a 2-D operand is shared)."""
    if _ag.needs_grad(a, b):
        return _ag.SptmMatmul.apply(a, b, trans_a, trans_b)
    a, b = _f64(a, "a"), _f64(b, "b")
    n = a.shape[-1]
    if a.shape[-2:] != (n, n) or b.shape[-2:] != (n, n):
        raise ValueError("sptm_matmul expects square matrices of equal size")
    batch = max(a.shape[0] if a.ndim == 3 else 1, b.shape[0] if b.ndim == 3 else 1)
    for t in (a, b):
        if t.ndim == 3 and t.shape[0] != batch:
            raise ValueError("batch sizes differ")
    sa = n * n if a.ndim == 3 else 0
    sb = n * n if b.ndim == 3 else 0
    out = torch.empty((batch, n, n), dtype=torch.float64, device=a.device)
    check(_lib.load().mcb200_sptm_matmul(_ptr(a), sa, int(trans_a), _ptr(b), sb, int(trans_b), _ptr(out), n * n, n,
                                         batch, _stream()))
    return out if (a.ndim == 3 or b.ndim == 3) else out[0]


def matchgate_sptm(u: torch.Tensor, atol: float = 1e-6) -> torch.Tensor:
    """(B, 4, 4) complex matchgates -> (B, 4, 4) real SPTM blocks; raises if the imaginary part exceeds ``atol``."""
    if not u.is_cuda:
        raise Mcb200Error("u must live on a CUDA device: matchcake_b200 has no CPU path")
    u = u.to(torch.complex128).contiguous()
    single = u.ndim == 2
    ur = torch.view_as_real(u.reshape(-1, 4, 4)).contiguous()
    B = ur.shape[0]
    out = torch.empty((B, 4, 4), dtype=torch.float64, device=u.device)
    worst = torch.zeros((1,), dtype=torch.float64, device=u.device)
    check(_lib.load().mcb200_matchgate_sptm(_ptr(ur), B, _ptr(out), _ptr(worst), _stream()))
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
    B, n = ar.shape[0], ar.shape[1]
    D = 2 * n + (1 if extended else 0)
    out = torch.empty((B, D, D), dtype=torch.float64, device=a.device)
    check(_lib.load().mcb200_product_state_cov(_ptr(ar), B, n, int(extended), _ptr(out), _stream()))
    return out[0] if single else out


# ------------------------------------------------------------------------------------------------ covariance
def cov_basis(x: torch.Tensor, n_qubits: int, bits: Optional[torch.Tensor] = None,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """cov[b] = X_b^T Lambda_0(bits) X_b for X (B, d, m); ``out`` (B, m, m) contiguous float64 receives the result."""
    if _ag.needs_grad(x):
        return _ag.CovBasis.apply(x, n_qubits, bits)
    x = _f64(x, "x")
    B, d, m = x.shape
    if d != 2 * n_qubits:
        raise ValueError("x must have 2*n_qubits rows")
    ib = None if bits is None else bits.to(device=x.device, dtype=torch.int8).contiguous()
    if out is None:
        out = torch.empty((B, m, m), dtype=torch.float64, device=x.device)
    elif out.shape != (B, m, m) or out.dtype != torch.float64 or not out.is_contiguous():
        raise ValueError(f"out must be a contiguous float64 ({B}, {m}, {m}) tensor")
    check(_lib.load().mcb200_cov_basis(_ptr(x), _ptr(ib), B, n_qubits, m, _ptr(out), _stream()))
    return out


def cov_dense(r: torch.Tensor, l0: torch.Tensor, idx: Optional[torch.Tensor] = None) -> torch.Tensor:
    """cov[b] = (R~_b^T L0 R~_b)[idx, idx]; L0 is (D,D) or (B,D,D) with D = d or d+1 (extended covariance)."""
    if _ag.needs_grad(l0):
        raise NotImplementedError("cov_dense is differentiable w.r.t. the SPTM only: the initial-state covariance is a constant")
    if _ag.needs_grad(r):
        return _ag.CovDense.apply(r, l0, idx)
    r, l0 = _f64(r, "r"), _f64(l0, "l0")
    if r.ndim == 2:
        r = r[None]
    Br, d, _ = r.shape
    D = l0.shape[-1]
    Bl = l0.shape[0] if l0.ndim == 3 else 1
    B = max(Br, Bl)
    if Br not in (1, B) or Bl not in (1, B):
        raise ValueError("batch sizes of r and l0 differ")
    sr = d * d if (Br == B and B > 1) else 0
    sl = D * D if (l0.ndim == 3 and Bl == B and B > 1) else 0
    if idx is not None:
        idx = idx.to(device=r.device, dtype=torch.int32).contiguous()
        m = idx.numel()
    else:
        m = D
    lib = _lib.load()
    ws = _workspace(lib.mcb200_cov_dense_workspace_bytes(B, D, m), r.device)
    out = torch.empty((B, m, m), dtype=torch.float64, device=r.device)
    check(lib.mcb200_cov_dense(_ptr(r), sr, _ptr(l0), sl, B, d, D, _ptr(idx), m, _ptr(out), _ptr(ws), _stream()))
    return out


# ------------------------------------------------------------------------------------------------ Pfaffians
def _apply_floor(v: torch.Tensor, floor2: float) -> torch.Tensor:
    """Differentiable restatement of the in-kernel floor for the autograd routes."""
    return torch.sqrt(v * v + floor2) if floor2 > 0.0 else v


def pfaffian(a: torch.Tensor, signed: bool = False, scale: float = 1.0, epsilon: float = 0.0) -> torch.Tensor:
    """Batched Pfaffian of real skew-symmetric matrices (..., m, m) -> (...).

    ``epsilon`` > 0 (magnitude mode only): ``scale * sqrt(Pf^2 + epsilon)``, the value of the reference's
    ``utils.pfaffian(sign=False, epsilon=...)`` = ``sqrt(|det A| + EPSILON)`` (utils/_pfaffian.py:110-118)."""
    eps = 0.0 if signed else float(epsilon)
    if _ag.needs_grad(a):
        return _apply_floor(_ag.Pfaffian.apply(a, signed, scale), scale * scale * eps)
    a = _f64(a, "a")
    m = a.shape[-1]
    if a.shape[-2] != m:
        raise ValueError("pfaffian expects square matrices")
    batch_shape = a.shape[:-2]
    n = int(np.prod(batch_shape)) if len(batch_shape) else 1
    out = torch.empty((n,), dtype=torch.float64, device=a.device)
    lib = _lib.load()
    ws = _workspace(lib.mcb200_pfaffian_workspace_bytes(n, m), a.device)
    with _Floor(scale * scale * eps):
        check(lib.mcb200_pfaffian(_ptr(a), n, m, PF_SIGNED if signed else PF_ABS, float(scale), _ptr(out), _ptr(ws),
                                  _stream()))
    return out.reshape(batch_shape)


def probs(cov: torch.Tensor, outcomes: torch.Tensor, normalize: bool = False, floor2: float = 0.0) -> torch.Tensor:
    """out[b, q] = 2^-k |Pf(cov[b] + Lambda_y(outcomes[q]))|; cov (B, 2k, 2k), outcomes (Q, k).

    ``floor2`` > 0 returns sqrt(p^2 + floor2) (before the optional normalisation): the reference's probabilities are
    ``sqrt(|det| + 1e-32)`` -- floor2 = 1e-32 for LookupTableStrategy (lookup_table_strategy.py:70-71), 4^-k 1e-32 for
    ProductStateProbabilityStrategy (product_state_strategy.py:211-212)."""
    if _ag.needs_grad(cov):
        p = _apply_floor(_ag.Probs.apply(cov, outcomes), floor2)
        return p / p.sum(dim=1, keepdim=True) if normalize else p
    cov = _f64(cov, "cov")
    outcomes = outcomes.to(device=cov.device, dtype=torch.int8).contiguous()
    B, m, _ = cov.shape
    Q, k = outcomes.shape
    if m != 2 * k:
        raise ValueError("cov must be (B, 2k, 2k) for outcomes (Q, k)")
    if m > 128:
        # more than 64 measured wires: the fused gather kernels stop at 128 x 128; form cov + Lambda_y explicitly (one
        # outcome at a time) and use the large-matrix Pfaffian kernels
        sgn = outcomes.to(torch.float64) * 2.0 - 1.0                       # (Q, k): +1 for bit 1, -1 for bit 0
        ev = torch.arange(0, m, 2, device=cov.device)
        out = torch.empty((B, Q), dtype=torch.float64, device=cov.device)
        for q in range(Q):
            mat = cov.clone()
            mat[:, ev, ev + 1] += sgn[q]
            mat[:, ev + 1, ev] -= sgn[q]
            out[:, q] = pfaffian(mat, signed=False, scale=2.0 ** -k, epsilon=floor2 * 4.0 ** k)
        return out / out.sum(dim=1, keepdim=True) if normalize else out
    out = torch.empty((B, Q), dtype=torch.float64, device=cov.device)
    with _Floor(floor2):
        check(_lib.load().mcb200_probs(_ptr(cov), _ptr(outcomes), B, Q, k, int(normalize), _ptr(out), _stream()))
    return out


def probs_all(cov: torch.Tensor, normalize: bool = False, floor2: float = 0.0) -> torch.Tensor:
    """Every outcome of the k measured wires: out[b, q] for q in range(2^k), MSB first; cov (B, 2k, 2k)."""
    if _ag.needs_grad(cov):  # the elimination tree has no adjoint; enumerate the outcomes (states_to_binary order)
        k = cov.shape[-1] // 2
        q = torch.arange(1 << k, device=cov.device)
        outcomes = ((q[:, None] >> torch.arange(k - 1, -1, -1, device=cov.device)[None, :]) & 1).to(torch.int8)
        return probs(cov, outcomes, normalize, floor2=floor2)
    cov = _f64(cov, "cov")
    B, m, _ = cov.shape
    k = m // 2
    if m != 2 * k or k < 1:
        raise ValueError("cov must be (B, 2k, 2k)")
    out = torch.empty((B, 1 << k), dtype=torch.float64, device=cov.device)
    lib = _lib.load()
    ws = _workspace(lib.mcb200_probs_all_workspace_bytes(k), cov.device)
    with _Floor(floor2):
        check(lib.mcb200_probs_all(_ptr(cov), B, k, int(normalize), _ptr(out), _ptr(ws), _stream()))
    return out


def sample(cov: torch.Tensor, shots: int, seed: int = 0, forced_bits: Optional[torch.Tensor] = None,
           return_prob: bool = False):
    """Draw ``shots`` bit strings per instance from p(y) = 2^-k |Pf(cov + Lambda_y)|; cov (B, 2k, 2k).

    Returns int8 ``(shots, B, k)`` (and the probability of every drawn string ``(shots, B)`` with ``return_prob``).
    With ``forced_bits`` (shots, B, k) nothing is drawn: the probabilities of the given strings are returned."""
    cov = _f64(cov.detach(), "cov")
    B, m, _ = cov.shape
    k = m // 2
    if m != 2 * k or k < 1:
        raise ValueError("cov must be (B, 2k, 2k)")
    fb = None
    if forced_bits is not None:
        fb = forced_bits.to(device=cov.device, dtype=torch.int8).contiguous()
        if tuple(fb.shape) != (shots, B, k):
            raise ValueError(f"forced_bits must have shape {(shots, B, k)}, got {tuple(fb.shape)}")
    bits = torch.empty((shots, B, k), dtype=torch.int8, device=cov.device) if fb is None else None
    prob = torch.empty((shots, B), dtype=torch.float64, device=cov.device) if (return_prob or fb is not None) else None
    check(_lib.load().mcb200_sample(_ptr(cov), B, k, shots, int(seed) & 0xFFFFFFFFFFFFFFFF, _ptr(fb), _ptr(bits),
                                    _ptr(prob), _stream()))
    if fb is not None:
        return prob
    return (bits, prob) if return_prob else bits


def sector_features(cov: torch.Tensor, index_sets: torch.Tensor) -> torch.Tensor:
    """feats[b, t] = Pf(cov[b][S_t, S_t]); cov (B, D, D), index_sets (T, s)."""
    if _ag.needs_grad(cov):
        return _ag.SectorFeatures.apply(cov, index_sets)
    cov = _f64(cov, "cov")
    index_sets = index_sets.to(device=cov.device, dtype=torch.int32).contiguous()
    B, D, _ = cov.shape
    T, s = index_sets.shape
    out = torch.empty((B, T), dtype=torch.float64, device=cov.device)
    check(_lib.load().mcb200_sector_features(_ptr(cov), B, D, _ptr(index_sets), T, s, _ptr(out), _stream()))
    return out


def rowdot(feats: torch.Tensor, coeffs: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[b] (+)= sum_t feats[b,t] coeffs[t] (accumulates when ``out`` is given)."""
    if _ag.needs_grad(feats, out):  # linear: autograd's own matvec is the adjoint
        r = feats.to(torch.float64) @ coeffs.to(device=feats.device, dtype=torch.float64)
        return r if out is None else out + r
    feats = _f64(feats, "feats")
    coeffs = _f64(coeffs.to(feats.device), "coeffs")
    B, T = feats.shape
    acc = out is not None
    if out is None:
        out = torch.empty((B,), dtype=torch.float64, device=feats.device)
    check(_lib.load().mcb200_rowdot(_ptr(feats), _ptr(coeffs), B, T, int(acc), _ptr(out), _stream()))
    return out


def pauli_sum_small(cov: torch.Tensor, term_ptr: torch.Tensor, indices: torch.Tensor, coeffs: torch.Tensor,
                    out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[b] (+)= sum_t coeffs[t] Pf(cov[b][S_t, S_t]) for index sets of mixed sizes <= 6 (one launch)."""
    if _ag.needs_grad(cov, out):
        raise NotImplementedError("pauli_sum_small is inference-only; use sector_features + rowdot under autograd")
    cov = _f64(cov, "cov")
    B, D = cov.shape[0], cov.shape[-1]
    term_ptr = term_ptr.to(device=cov.device, dtype=torch.int32).contiguous()
    indices = indices.to(device=cov.device, dtype=torch.int32).contiguous()
    coeffs = _f64(coeffs.to(cov.device), "coeffs")
    T = coeffs.shape[0]
    if term_ptr.shape[0] != T + 1:
        raise Mcb200Error(f"pauli_sum_small: term_ptr must have {T + 1} entries, got {term_ptr.shape[0]}")
    acc = out is not None
    if out is None:
        out = torch.empty((B,), dtype=torch.float64, device=cov.device)
    check(_lib.load().mcb200_pauli_sum_small(_ptr(cov), B, D, _ptr(term_ptr), _ptr(indices), _ptr(coeffs), T, int(acc),
                                             _ptr(out), _stream()))
    return out


# ------------------------------------------------------------------------------------------------ Gram
def _gram_differentiable(cov0, cov1, scale, mode, row_range, floor2=0.0):
    """Training-size Gram under autograd (kernel alignment, ml/kernels/kernel.py:200-283): the evaluated cells are
    batched Pfaffians of cov0[i] + cov1[j] through the differentiable Pfaffian op; cell selection, mirroring and the
    unit diagonal follow gram_matrix.py:127-161,191-195."""
    n0, n1 = cov0.shape[0], cov1.shape[0]
    rb, re = (0, n0) if row_range is None else row_range
    dev = cov0.device
    ii, jj = torch.meshgrid(torch.arange(rb, re, device=dev), torch.arange(n1, device=dev), indexing="ij")
    if mode == GRAM_FULL:
        keep = torch.ones_like(ii, dtype=torch.bool)
    else:
        keep = (jj > ii) if n0 < n1 else (ii > jj)
    ii, jj = ii[keep], jj[keep]
    vals = _apply_floor(pfaffian(cov0[ii] + cov1[jj], signed=False, scale=scale), floor2)
    K = torch.zeros((n0, n1), dtype=torch.float64, device=dev).index_put((ii, jj), vals)
    if mode == GRAM_REFERENCE:
        q = min(n0, n1)
        mirror = (jj < n0) & (ii < n1)
        K = K.index_put((jj[mirror], ii[mirror]), vals[mirror])
        eye = torch.zeros_like(K)
        eye[torch.arange(q, device=dev), torch.arange(q, device=dev)] = 1.0
        K = K + eye
    return K


def _gram_out(out: Optional[torch.Tensor], n0: int, n1: int, rb: int, re: int, rows_only: bool, device):
    """Output buffer of a Gram launch and the pointer the kernel sees as ``K`` (row 0 of the FULL matrix).

    ``rows_only``: ``out`` holds just the rows [rb, re) -- shape (re - rb, n1), e.g. a view into the final matrix of a
    row-sharded build; the kernel only writes ``K[i * ldk + j]`` for rb <= i < re (mode TRIANGLE / FULL), so ``K`` is
    the buffer shifted back by ``rb`` rows and nothing of size (n0, n1) is ever allocated or zero-filled."""
    if rows_only:
        if out is None:
            out = torch.empty((re - rb, n1), dtype=torch.float64, device=device)
        if out.shape != (re - rb, n1) or out.dtype != torch.float64 or out.stride(1) != 1:
            raise ValueError(f"rows-only Gram buffer must be float64 ({re - rb}, {n1}) with unit column stride")
        return out, out.data_ptr() - rb * out.stride(0) * 8 if re > rb else out.data_ptr()
    if out is None:
        out = torch.zeros((n0, n1), dtype=torch.float64, device=device)
    return out, out.data_ptr()


def gram(cov0: torch.Tensor, cov1: torch.Tensor, scale: float, mode: int = GRAM_REFERENCE,
         row_range: Optional[Tuple[int, int]] = None, out: Optional[torch.Tensor] = None,
         rows_only: bool = False, floor2: float = 0.0) -> torch.Tensor:
    """K[i, j] = scale |Pf(cov0[i] + cov1[j])| with the reference's triangle/mirror/unit-diagonal rules.
    ``floor2`` > 0: evaluated entries are sqrt(K^2 + floor2) (the reference's Gram goes through LookupTableStrategy,
    i.e. sqrt(|det| + 1e-32), nif_kernel.py:210-221 -> lookup_table_strategy.py:70-71)."""
    if _ag.needs_grad(cov0, cov1):
        return _gram_differentiable(cov0, cov1, scale, mode, row_range, floor2)
    cov0, cov1 = _f64(cov0, "cov0"), _f64(cov1, "cov1")
    n0, d, _ = cov0.shape
    n1 = cov1.shape[0]
    if cov1.shape[1:] != (d, d):
        raise ValueError("cov0 and cov1 must hold matrices of the same size")
    rb, re = (0, n0) if row_range is None else row_range
    if rows_only and mode == GRAM_REFERENCE:
        raise ValueError("rows_only needs mode TRIANGLE or FULL (REFERENCE mirrors into rows outside the block)")
    out, kptr = _gram_out(out, n0, n1, rb, re, rows_only, cov0.device)
    lib = _lib.load()
    ws = _workspace(lib.mcb200_gram_workspace_bytes(n0, n1, d), cov0.device)
    with _Floor(floor2):
        check(lib.mcb200_gram(_ptr(cov0), _ptr(cov1), n0, n1, d, rb, re, mode, float(scale), kptr, out.stride(0),
                              _ptr(ws), _stream()))
    return out


def circuit_pair_blocks(circ: CompiledCircuit, params: torch.Tensor, pair_wire0: torch.Tensor,
                        init_bits: Optional[torch.Tensor] = None) -> torch.Tensor:
    """(B, C, 6) covariance blocks of a circuit whose wire pairs (w, w+1), w in ``pair_wire0``, never interact."""
    params = circ._params(params)
    B = params.shape[0]
    pw = pair_wire0.to(device=params.device, dtype=torch.int32).contiguous()
    ib = None if init_bits is None else init_bits.to(device=params.device, dtype=torch.int8).contiguous()
    out = torch.empty((B, pw.numel(), 6), dtype=torch.float64, device=params.device)
    check(_lib.load().mcb200_circuit_pair_blocks(C.byref(circ._struct), _ptr(params), B, _ptr(pw), pw.numel(), _ptr(ib),
                                                 _ptr(out), _stream()))
    return out


def gram_blocks4(blocks0: torch.Tensor, blocks1: torch.Tensor, mode: int = GRAM_REFERENCE,
                 row_range: Optional[Tuple[int, int]] = None, out: Optional[torch.Tensor] = None,
                 rows_only: bool = False, floor2: float = 0.0) -> torch.Tensor:
    """Gram of block-diagonal covariances: blocks (n, C, 6) = upper triangles of the C 4x4 diagonal blocks."""
    blocks0, blocks1 = _f64(blocks0, "blocks0"), _f64(blocks1, "blocks1")
    n0, C, six = blocks0.shape
    n1 = blocks1.shape[0]
    if six != 6 or blocks1.shape[1:] != (C, 6):
        raise ValueError("blocks must be (n, C, 6)")
    rb, re = (0, n0) if row_range is None else row_range
    if rows_only and mode == GRAM_REFERENCE:
        raise ValueError("rows_only needs mode TRIANGLE or FULL (REFERENCE mirrors into rows outside the block)")
    out, kptr = _gram_out(out, n0, n1, rb, re, rows_only, blocks0.device)
    with _Floor(floor2):
        check(_lib.load().mcb200_gram_blocks4(_ptr(blocks0), _ptr(blocks1), n0, n1, C, rb, re, mode, 2.0 ** (-2 * C),
                                              kptr, out.stride(0), _stream()))
    return out


def gram_symmetrize(k: torch.Tensor) -> torch.Tensor:
    check(_lib.load().mcb200_gram_symmetrize(_ptr(k), k.shape[0], k.shape[1], k.stride(0), _stream()))
    return k


def probe_fp64_tflops(kind: int, iters: int = 2000) -> float:
    v = _lib.load().mcb200_probe_fp64_tflops(kind, iters, _stream())
    if v < 0:
        check(1)
    return float(v)
