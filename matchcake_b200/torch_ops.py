"""``torch.ops.mcb200.*`` -- the C ABI (include/mcb200.h) registered as torch custom operators.

BASELINE.json's north_star asks for "a thin C-ABI torch custom op in Python" (last row of the
boundary table).  Every compute entry point of libmcb200.so is registered here through ``torch.library``:

  * ``torch.library.define("mcb200::<op>", schema)``        -- the operator and its schema,
  * ``torch.library.impl(..., "CUDA")``                     -- the kernel: a ctypes call into libmcb200.so with raw
    device pointers and the current torch stream (no torch types cross the ABI),
  * ``torch.library.register_fake``                          -- shape/dtype propagation, so FakeTensor / ``meta`` tracing
    (``torch.compile``, ``torch.export``, ``make_fx``) see through the ops without launching anything,
  * ``torch.library.register_autograd``                      -- reverse mode: the hand-written adjoint kernels of
    ``csrc/backward.cu`` (themselves registered as ``mcb200::*_backward`` ops), replacing the reference's torch
    autograd over its op sequence (devices/nif_device.py:949-989, ml/kernels/kernel.py:200-283).

The low-level ``define`` / ``impl`` pair is used instead of the ``torch.library.custom_op`` decorator because the
decorator's per-call Python wrapper costs ~0.5 ms (measured, torch 2.11) against ~10 us here -- the fused C2 step is
0.17 ms.  There is only a CUDA kernel: a CPU tensor fails in the dispatcher ("no kernel for backend CPU"), and a
missing library raises in ``_lib.load`` -- no fallback of any kind.

Gate-list circuits are not tensors; ``ops.CompiledCircuit`` registers itself here and the ops take its integer
``handle`` (weakly referenced: the circuit object owns the device descriptor).
"""
from __future__ import annotations

import ctypes as C
import itertools
import weakref
from typing import Optional, Tuple

import torch
from torch import library as _tl

from . import _lib
from ._lib import Mcb200Error, check

NS = "mcb200"
PF_ABS, PF_SIGNED = 0, 1
_U64 = 0xFFFFFFFFFFFFFFFF

# ------------------------------------------------------------------------------------------------ circuit handles
_circuits: "weakref.WeakValueDictionary[int, object]" = weakref.WeakValueDictionary()
_next_handle = itertools.count(1)


def register_circuit(circ) -> int:
    h = next(_next_handle)
    _circuits[h] = circ
    return h


def circuit(handle: int):
    try:
        return _circuits[int(handle)]
    except KeyError:
        raise Mcb200Error(f"circuit handle {handle} is not alive (the CompiledCircuit object was released)") from None


# ------------------------------------------------------------------------------------------------ helpers
_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream() -> int:
    """cudaStream_t of torch's current stream on the current device (the raw getter is ~10x cheaper than building a
    ``torch.cuda.Stream`` object per launch)."""
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _d(t: torch.Tensor, name: str) -> torch.Tensor:
    """float64 + contiguous, enforced (the kernels read raw pointers)."""
    if t.dtype != torch.float64:
        raise Mcb200Error(f"mcb200: {name} must be float64, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


def _i(t: Optional[torch.Tensor], dtype, name: str) -> Optional[torch.Tensor]:
    if t is None:
        return None
    if t.dtype != dtype:
        raise Mcb200Error(f"mcb200: {name} must be {dtype}, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


def _workspace(nbytes: int, device) -> Optional[torch.Tensor]:
    if nbytes <= 0:
        return None
    return torch.empty((nbytes + 7) // 8, dtype=torch.float64, device=device)


class _Floor:
    """``with _Floor(floor2):`` -- magnitude floor of the launches inside the block (include/mcb200.h:
    mcb200_set_abs_floor; the option is per thread and always restored to 0 = off)."""

    __slots__ = ("floor2",)

    def __init__(self, floor2: float):
        self.floor2 = float(floor2)

    def __enter__(self):
        if self.floor2 != 0.0:
            check(_lib.load().mcb200_set_abs_floor(self.floor2))

    def __exit__(self, *exc):
        if self.floor2 != 0.0:
            _lib.load().mcb200_set_abs_floor(0.0)
        return False


OPS = {}  # name -> schema (tests/test_library_abi.py checks that every compute entry point has an operator)


def _define(name: str, schema: str):
    _tl.define(f"{NS}::{name}", schema)
    OPS[name] = schema

    def deco(fn):
        _tl.impl(f"{NS}::{name}", "CUDA")(fn)
        return fn

    return deco


def _fake(name: str):
    return _tl.register_fake(f"{NS}::{name}")


def _no_floor_grad(floor2: float, what: str):
    if floor2 != 0.0:
        raise NotImplementedError(f"mcb200::{what}: the in-kernel magnitude floor has no adjoint; call with floor2 = 0 "
                                  "and apply sqrt(v^2 + floor2) in torch (matchcake_b200.ops does)")


# ------------------------------------------------------------------------------------------------ K1: chain
@_define("circuit_columns", "(Tensor params, int circuit, Tensor? cols) -> Tensor")
def _circuit_columns(params, circuit_handle, cols):
    circ = circuit(circuit_handle)
    params = _d(params, "params")
    cols = _i(cols, torch.int32, "cols")
    B = params.shape[0]
    m = circ.d if cols is None else cols.numel()
    out = torch.empty((B, circ.d, m), dtype=torch.float64, device=params.device)
    check(_lib.load().mcb200_circuit_columns(C.byref(circ._struct), _p(params), B, _p(cols), m, _p(out), _stream()))
    return out


@_fake("circuit_columns")
def _(params, circuit_handle, cols):
    circ = circuit(circuit_handle)
    m = circ.d if cols is None else cols.numel()
    return params.new_empty((params.shape[0], circ.d, m), dtype=torch.float64)


@_define("circuit_columns_backward", "(Tensor params, int circuit, Tensor x, Tensor gx) -> Tensor")
def _circuit_columns_backward(params, circuit_handle, x, gx):
    circ = circuit(circuit_handle)
    params, x, gx = _d(params, "params"), _d(x, "x"), _d(gx, "gx")
    B, _, m = x.shape
    ga = torch.empty((B, circ.n_param_cols), dtype=torch.float64, device=x.device)
    check(_lib.load().mcb200_circuit_columns_backward(C.byref(circ._struct), _p(params), B, m, _p(x), _p(gx), _p(ga),
                                                      _stream()))
    return ga


@_fake("circuit_columns_backward")
def _(params, circuit_handle, x, gx):
    return params.new_empty(params.shape, dtype=torch.float64)


def _cc_setup(ctx, inputs, output):
    params, handle, _cols = inputs
    ctx.handle = handle
    ctx.save_for_backward(params.detach(), output.detach())


def _cc_backward(ctx, gx):
    params, x = ctx.saved_tensors
    ga = torch.ops.mcb200.circuit_columns_backward(params, ctx.handle, x, gx.to(torch.float64).contiguous())
    circ = circuit(ctx.handle)
    if circ.scale_dev is not None:  # angle = bias + scale * x was applied in-kernel
        ga = ga * circ.scale_dev[None, :]
    return ga, None, None


_tl.register_autograd(f"{NS}::circuit_columns", _cc_backward, setup_context=_cc_setup)


@_define("circuit_cov_basis", "(Tensor params, int circuit, Tensor cols, Tensor? bits) -> Tensor")
def _circuit_cov_basis(params, circuit_handle, cols, bits):
    """cov[b] = (R_b^T Lambda_0(bits) R_b)[cols, cols] straight from the gate list (K1 + K2 fused when the panel fits)."""
    circ = circuit(circuit_handle)
    params = _d(params, "params")
    cols, ib = _i(cols, torch.int32, "cols"), _i(bits, torch.int8, "bits")
    B, m = params.shape[0], cols.numel()
    lib = _lib.load()
    ws = _workspace(lib.mcb200_circuit_cov_basis_workspace_bytes(C.byref(circ._struct), B, m), params.device)
    out = torch.empty((B, m, m), dtype=torch.float64, device=params.device)
    check(lib.mcb200_circuit_cov_basis(C.byref(circ._struct), _p(params), B, _p(cols), m, _p(ib), _p(out), _p(ws),
                                       _stream()))
    return out


@_fake("circuit_cov_basis")
def _(params, circuit_handle, cols, bits):
    return params.new_empty((params.shape[0], cols.numel(), cols.numel()), dtype=torch.float64)


@_define("circuit_expval_projector",
         "(Tensor params, int circuit, Tensor? init_bits, Tensor? target_bits, float floor2) -> Tensor")
def _circuit_expval_projector(params, circuit_handle, init_bits, target_bits, floor2):
    circ = circuit(circuit_handle)
    params = _d(params, "params")
    ib, tb = _i(init_bits, torch.int8, "init_bits"), _i(target_bits, torch.int8, "target_bits")
    B = params.shape[0]
    out = torch.empty((B,), dtype=torch.float64, device=params.device)
    with _Floor(floor2):
        check(_lib.load().mcb200_circuit_expval_projector(C.byref(circ._struct), _p(params), B, _p(ib), _p(tb),
                                                          _p(out), _stream()))
    return out


@_fake("circuit_expval_projector")
def _(params, circuit_handle, init_bits, target_bits, floor2):
    return params.new_empty((params.shape[0],), dtype=torch.float64)


@_define("circuit_pair_blocks", "(Tensor params, int circuit, Tensor pair_wire0, Tensor? init_bits) -> Tensor")
def _circuit_pair_blocks(params, circuit_handle, pair_wire0, init_bits):
    circ = circuit(circuit_handle)
    params = _d(params, "params")
    pw, ib = _i(pair_wire0, torch.int32, "pair_wire0"), _i(init_bits, torch.int8, "init_bits")
    B = params.shape[0]
    out = torch.empty((B, pw.numel(), 6), dtype=torch.float64, device=params.device)
    check(_lib.load().mcb200_circuit_pair_blocks(C.byref(circ._struct), _p(params), B, _p(pw), pw.numel(), _p(ib),
                                                 _p(out), _stream()))
    return out


@_fake("circuit_pair_blocks")
def _(params, circuit_handle, pair_wire0, init_bits):
    return params.new_empty((params.shape[0], pair_wire0.numel(), 6), dtype=torch.float64)


@_define("sptm_matmul", "(Tensor a, Tensor b, bool trans_a, bool trans_b) -> Tensor")
def _sptm_matmul(a, b, trans_a, trans_b):
    """(n,n) or (B,n,n) operands (a 2-D operand is shared by the batch) -> (B,n,n)."""
    a, b = _d(a, "a"), _d(b, "b")
    n = a.shape[-1]
    batch = max(a.shape[0] if a.ndim == 3 else 1, b.shape[0] if b.ndim == 3 else 1)
    sa = n * n if a.ndim == 3 else 0
    sb = n * n if b.ndim == 3 else 0
    out = torch.empty((batch, n, n), dtype=torch.float64, device=a.device)
    check(_lib.load().mcb200_sptm_matmul(_p(a), sa, int(trans_a), _p(b), sb, int(trans_b), _p(out), n * n, n, batch,
                                         _stream()))
    return out


@_fake("sptm_matmul")
def _(a, b, trans_a, trans_b):
    n = a.shape[-1]
    batch = max(a.shape[0] if a.ndim == 3 else 1, b.shape[0] if b.ndim == 3 else 1)
    return a.new_empty((batch, n, n), dtype=torch.float64)


def _mm_setup(ctx, inputs, output):
    a, b, ta, tb = inputs
    ctx.trans = (ta, tb)
    ctx.save_for_backward(a.detach(), b.detach())


def _mm_backward(ctx, gc):
    """C = A' B' with A' = op(a), B' = op(b):  gA' = gC B'^T,  gB' = A'^T gC -- two more chain steps on the same
    DMMA kernel."""
    mm = torch.ops.mcb200.sptm_matmul
    a, b = ctx.saved_tensors
    ta, tb = ctx.trans
    gc = gc.to(torch.float64).contiguous()
    ga = mm(gc, b, False, not tb) if not ta else mm(b, gc, tb, True)
    gb = mm(a, gc, not ta, False) if not tb else mm(gc, a, True, ta)
    if a.ndim == 2:  # a shared operand receives the sum over the batch
        ga = ga.sum(0)
    if b.ndim == 2:
        gb = gb.sum(0)
    return ga, gb, None, None


_tl.register_autograd(f"{NS}::sptm_matmul", _mm_backward, setup_context=_mm_setup)


@_define("matchgate_sptm", "(Tensor u_re_im) -> (Tensor, Tensor)")
def _matchgate_sptm(u_re_im):
    """(B, 4, 4, 2) interleaved complex matchgates -> ((B, 4, 4) real SPTM blocks, (1,) max |Im|)."""
    ur = _d(u_re_im, "u_re_im")
    B = ur.shape[0]
    out = torch.empty((B, 4, 4), dtype=torch.float64, device=ur.device)
    worst = torch.zeros((1,), dtype=torch.float64, device=ur.device)
    check(_lib.load().mcb200_matchgate_sptm(_p(ur), B, _p(out), _p(worst), _stream()))
    return out, worst


@_fake("matchgate_sptm")
def _(u_re_im):
    return u_re_im.new_empty((u_re_im.shape[0], 4, 4), dtype=torch.float64), u_re_im.new_empty((1,), dtype=torch.float64)


@_define("matchgate_sptm_backward", "(Tensor u_re_im, Tensor grad_blocks) -> Tensor")
def _matchgate_sptm_backward(u_re_im, grad_blocks):
    ur, g = _d(u_re_im, "u_re_im"), _d(grad_blocks, "grad_blocks")
    gu = torch.empty_like(ur)
    check(_lib.load().mcb200_matchgate_sptm_backward(_p(ur), ur.shape[0], _p(g), _p(gu), _stream()))
    return gu


@_fake("matchgate_sptm_backward")
def _(u_re_im, grad_blocks):
    return u_re_im.new_empty(u_re_im.shape, dtype=torch.float64)


def _mg_setup(ctx, inputs, output):
    ctx.save_for_backward(inputs[0].detach())
    ctx.set_materialize_grads(False)


def _mg_backward(ctx, g_blocks, g_worst):
    (ur,) = ctx.saved_tensors
    if g_blocks is None:
        return None
    return torch.ops.mcb200.matchgate_sptm_backward(ur, g_blocks.to(torch.float64).contiguous())


_tl.register_autograd(f"{NS}::matchgate_sptm", _mg_backward, setup_context=_mg_setup)


@_define("product_state_cov", "(Tensor amplitudes_re_im, bool extended) -> Tensor")
def _product_state_cov(amp, extended):
    """(B, n, 2, 2) interleaved complex per-qubit amplitudes -> (B, D, D), D = 2n (+1 if extended)."""
    ar = _d(amp, "amplitudes_re_im")
    B, n = ar.shape[0], ar.shape[1]
    D = 2 * n + (1 if extended else 0)
    out = torch.empty((B, D, D), dtype=torch.float64, device=ar.device)
    check(_lib.load().mcb200_product_state_cov(_p(ar), B, n, int(extended), _p(out), _stream()))
    return out


@_fake("product_state_cov")
def _(amp, extended):
    D = 2 * amp.shape[1] + (1 if extended else 0)
    return amp.new_empty((amp.shape[0], D, D), dtype=torch.float64)


# ------------------------------------------------------------------------------------------------ K2: covariance
def _cov_basis_launch(out, x, n_qubits, bits):
    x = _d(x, "x")
    B, d, m = x.shape
    if d != 2 * n_qubits:
        raise Mcb200Error("mcb200::cov_basis: x must have 2*n_qubits rows")
    ib = _i(bits, torch.int8, "bits")
    check(_lib.load().mcb200_cov_basis(_p(x), _p(ib), B, n_qubits, m, _p(out), _stream()))


@_define("cov_basis", "(Tensor x, int n_qubits, Tensor? bits) -> Tensor")
def _cov_basis(x, n_qubits, bits):
    B, _, m = x.shape
    out = torch.empty((B, m, m), dtype=torch.float64, device=x.device)
    _cov_basis_launch(out, x, n_qubits, bits)
    return out


@_fake("cov_basis")
def _(x, n_qubits, bits):
    return x.new_empty((x.shape[0], x.shape[2], x.shape[2]), dtype=torch.float64)


@_define("cov_basis_out", "(Tensor(a!) out, Tensor x, int n_qubits, Tensor? bits) -> ()")
def _cov_basis_out(out, x, n_qubits, bits):
    B, _, m = x.shape
    if out.shape != (B, m, m) or out.dtype != torch.float64 or not out.is_contiguous():
        raise ValueError(f"out must be a contiguous float64 ({B}, {m}, {m}) tensor")
    _cov_basis_launch(out, x, n_qubits, bits)


@_fake("cov_basis_out")
def _(out, x, n_qubits, bits):
    return None


@_define("cov_basis_backward", "(Tensor x, int n_qubits, Tensor? bits, Tensor gcov) -> Tensor")
def _cov_basis_backward(x, n_qubits, bits, gcov):
    x, gcov = _d(x, "x"), _d(gcov, "gcov")
    B, _, m = x.shape
    ib = _i(bits, torch.int8, "bits")
    gx = torch.empty_like(x)
    check(_lib.load().mcb200_cov_basis_backward(_p(x), _p(ib), B, n_qubits, m, _p(gcov), _p(gx), _stream()))
    return gx


@_fake("cov_basis_backward")
def _(x, n_qubits, bits, gcov):
    return x.new_empty(x.shape, dtype=torch.float64)


def _cb_setup(ctx, inputs, output):
    x, n_qubits, bits = inputs
    ctx.n_qubits, ctx.bits = n_qubits, bits
    ctx.save_for_backward(x.detach())


def _cb_backward(ctx, gcov):
    (x,) = ctx.saved_tensors
    gx = torch.ops.mcb200.cov_basis_backward(x, ctx.n_qubits, ctx.bits, gcov.to(torch.float64).contiguous())
    return gx, None, None


_tl.register_autograd(f"{NS}::cov_basis", _cb_backward, setup_context=_cb_setup)


def _cov_dense_dims(r, l0):
    Br = r.shape[0] if r.ndim == 3 else 1
    d = r.shape[-1]
    D = l0.shape[-1]
    Bl = l0.shape[0] if l0.ndim == 3 else 1
    B = max(Br, Bl)
    if Br not in (1, B) or Bl not in (1, B):
        raise ValueError("batch sizes of r and l0 differ")
    sr = d * d if (Br == B and B > 1) else 0
    sl = D * D if (l0.ndim == 3 and Bl == B and B > 1) else 0
    return B, d, D, sr, sl


@_define("cov_dense", "(Tensor r, Tensor l0, Tensor? idx) -> Tensor")
def _cov_dense(r, l0, idx):
    r, l0 = _d(r, "r"), _d(l0, "l0")
    B, d, D, sr, sl = _cov_dense_dims(r, l0)
    idx = _i(idx, torch.int32, "idx")
    m = D if idx is None else idx.numel()
    lib = _lib.load()
    ws = _workspace(lib.mcb200_cov_dense_workspace_bytes(B, D, m), r.device)
    out = torch.empty((B, m, m), dtype=torch.float64, device=r.device)
    check(lib.mcb200_cov_dense(_p(r), sr, _p(l0), sl, B, d, D, _p(idx), m, _p(out), _p(ws), _stream()))
    return out


@_fake("cov_dense")
def _(r, l0, idx):
    B = max(r.shape[0] if r.ndim == 3 else 1, l0.shape[0] if l0.ndim == 3 else 1)
    m = l0.shape[-1] if idx is None else idx.numel()
    return r.new_empty((B, m, m), dtype=torch.float64)


@_define("cov_dense_backward", "(Tensor r, Tensor l0, Tensor? idx, Tensor gcov) -> Tensor")
def _cov_dense_backward(r, l0, idx, gcov):
    """Adjoint w.r.t. R only (the initial-state covariance is a constant); a shared R (2-D or batch of one) receives
    the sum over the batch (the kernel accumulates it when strideR == 0).  Returns (B or 1, d, d)."""
    r, l0, gcov = _d(r, "r"), _d(l0, "l0"), _d(gcov, "gcov")
    _, d, D, _, _ = _cov_dense_dims(r, l0)
    B, m, _ = gcov.shape
    Br = r.shape[0] if r.ndim == 3 else 1
    sr = d * d if (Br == B and B > 1) else 0
    sl = D * D if (l0.ndim == 3 and l0.shape[0] == B and B > 1) else 0
    idx = _i(idx, torch.int32, "idx")
    lib = _lib.load()
    ws = _workspace(lib.mcb200_cov_dense_backward_workspace_bytes(B, D, m), r.device)
    gr = torch.empty((B if sr else 1, d, d), dtype=torch.float64, device=r.device)
    check(lib.mcb200_cov_dense_backward(_p(r), sr, _p(l0), sl, B, d, D, _p(idx), m, _p(gcov), _p(gr), _p(ws),
                                        _stream()))
    return gr


@_fake("cov_dense_backward")
def _(r, l0, idx, gcov):
    B = gcov.shape[0]
    Br = r.shape[0] if r.ndim == 3 else 1
    d = r.shape[-1]
    return r.new_empty((B if (Br == B and B > 1) else 1, d, d), dtype=torch.float64)


def _cd_setup(ctx, inputs, output):
    r, l0, idx = inputs
    ctx.idx = idx
    ctx.r_shape = tuple(r.shape)
    ctx.save_for_backward(r.detach(), l0.detach())


def _cd_backward(ctx, gcov):
    r, l0 = ctx.saved_tensors
    gr = torch.ops.mcb200.cov_dense_backward(r, l0, ctx.idx, gcov.to(torch.float64).contiguous())
    return gr.reshape(ctx.r_shape), None, None


_tl.register_autograd(f"{NS}::cov_dense", _cd_backward, setup_context=_cd_setup)


# ------------------------------------------------------------------------------------------------ K3: Pfaffians
@_define("pfaffian", "(Tensor a, bool signed, float scale, float floor2) -> Tensor")
def _pfaffian(a, signed, scale, floor2):
    """a (n, m, m) real skew-symmetric -> (n,): scale Pf (signed) or sqrt((scale |Pf|)^2 + floor2)."""
    a = _d(a, "a")
    n, m = a.shape[0], a.shape[-1]
    out = torch.empty((n,), dtype=torch.float64, device=a.device)
    lib = _lib.load()
    ws = _workspace(lib.mcb200_pfaffian_workspace_bytes(n, m), a.device)
    with _Floor(0.0 if signed else floor2):
        check(lib.mcb200_pfaffian(_p(a), n, m, PF_SIGNED if signed else PF_ABS, float(scale), _p(out), _p(ws),
                                  _stream()))
    return out


@_fake("pfaffian")
def _(a, signed, scale, floor2):
    return a.new_empty((a.shape[0],), dtype=torch.float64)


@_define("pfaffian_backward", "(Tensor a, Tensor out, Tensor gout) -> Tensor")
def _pfaffian_backward(a, out, gout):
    a, out, gout = _d(a, "a"), _d(out, "out"), _d(gout, "gout")
    n, m = out.numel(), a.shape[-1]
    ga = torch.empty_like(a)
    if n and m:
        check(_lib.load().mcb200_pfaffian_backward(_p(a), n, m, _p(out), _p(gout), _p(ga), _stream()))
    return ga


@_fake("pfaffian_backward")
def _(a, out, gout):
    return a.new_empty(a.shape, dtype=torch.float64)


def _pf_setup(ctx, inputs, output):
    a, _signed, _scale, floor2 = inputs
    ctx.floor2 = floor2
    ctx.save_for_backward(a.detach(), output.detach())


def _pf_backward(ctx, gout):
    _no_floor_grad(ctx.floor2, "pfaffian")
    a, out = ctx.saved_tensors
    return torch.ops.mcb200.pfaffian_backward(a, out, gout.to(torch.float64).contiguous()), None, None, None


_tl.register_autograd(f"{NS}::pfaffian", _pf_backward, setup_context=_pf_setup)


@_define("probs", "(Tensor cov, Tensor outcomes, bool normalize, float floor2) -> Tensor")
def _probs(cov, outcomes, normalize, floor2):
    cov = _d(cov, "cov")
    outcomes = _i(outcomes, torch.int8, "outcomes")
    B, m, _ = cov.shape
    Q, k = outcomes.shape
    if m != 2 * k:
        raise ValueError("cov must be (B, 2k, 2k) for outcomes (Q, k)")
    out = torch.empty((B, Q), dtype=torch.float64, device=cov.device)
    with _Floor(floor2):
        check(_lib.load().mcb200_probs(_p(cov), _p(outcomes), B, Q, k, int(normalize), _p(out), _stream()))
    return out


@_fake("probs")
def _(cov, outcomes, normalize, floor2):
    return cov.new_empty((cov.shape[0], outcomes.shape[0]), dtype=torch.float64)


@_define("probs_backward", "(Tensor cov, Tensor outcomes, Tensor p, Tensor gp) -> Tensor")
def _probs_backward(cov, outcomes, p, gp):
    cov, p, gp = _d(cov, "cov"), _d(p, "p"), _d(gp, "gp")
    outcomes = _i(outcomes, torch.int8, "outcomes")
    B, Q = p.shape
    k = cov.shape[-1] // 2
    gcov = torch.empty_like(cov)
    check(_lib.load().mcb200_probs_backward(_p(cov), _p(outcomes), B, Q, k, _p(p), _p(gp), _p(gcov), _stream()))
    return gcov


@_fake("probs_backward")
def _(cov, outcomes, p, gp):
    return cov.new_empty(cov.shape, dtype=torch.float64)


def _pr_setup(ctx, inputs, output):
    cov, outcomes, normalize, floor2 = inputs
    ctx.outcomes, ctx.normalize, ctx.floor2 = outcomes, normalize, floor2
    ctx.save_for_backward(cov.detach(), output.detach())


def _pr_backward(ctx, gp):
    _no_floor_grad(ctx.floor2, "probs")
    if ctx.normalize:
        raise NotImplementedError("mcb200::probs: differentiate the unnormalised probabilities and normalise in torch")
    cov, p = ctx.saved_tensors
    return torch.ops.mcb200.probs_backward(cov, ctx.outcomes, p, gp.to(torch.float64).contiguous()), None, None, None


_tl.register_autograd(f"{NS}::probs", _pr_backward, setup_context=_pr_setup)


@_define("probs_all", "(Tensor cov, bool normalize, float floor2) -> Tensor")
def _probs_all(cov, normalize, floor2):
    cov = _d(cov, "cov")
    B, m, _ = cov.shape
    k = m // 2
    if m != 2 * k or k < 1:
        raise ValueError("cov must be (B, 2k, 2k)")
    out = torch.empty((B, 1 << k), dtype=torch.float64, device=cov.device)
    lib = _lib.load()
    ws = _workspace(lib.mcb200_probs_all_workspace_bytes(k), cov.device)
    with _Floor(floor2):
        check(lib.mcb200_probs_all(_p(cov), B, k, int(normalize), _p(out), _p(ws), _stream()))
    return out


@_fake("probs_all")
def _(cov, normalize, floor2):
    return cov.new_empty((cov.shape[0], 1 << (cov.shape[-1] // 2)), dtype=torch.float64)


@_define("sample", "(Tensor cov, int shots, int seed, Tensor? forced_bits, bool want_prob) -> (Tensor, Tensor)")
def _sample(cov, shots, seed, forced_bits, want_prob):
    """Returns (bits int8 (shots, B, k), prob (shots, B)); an output that was not asked for is an empty tensor
    (bits when ``forced_bits`` is given, prob unless ``want_prob`` or ``forced_bits``)."""
    cov = _d(cov, "cov")
    B, m, _ = cov.shape
    k = m // 2
    if m != 2 * k or k < 1:
        raise ValueError("cov must be (B, 2k, 2k)")
    fb = _i(forced_bits, torch.int8, "forced_bits")
    if fb is not None and tuple(fb.shape) != (shots, B, k):
        raise ValueError(f"forced_bits must have shape {(shots, B, k)}, got {tuple(fb.shape)}")
    bits = torch.empty((shots, B, k) if fb is None else (0,), dtype=torch.int8, device=cov.device)
    need_p = want_prob or fb is not None
    prob = torch.empty((shots, B) if need_p else (0,), dtype=torch.float64, device=cov.device)
    check(_lib.load().mcb200_sample(_p(cov), B, k, shots, int(seed) & _U64, _p(fb), _p(bits) if fb is None else None,
                                    _p(prob) if need_p else None, _stream()))
    return bits, prob


@_fake("sample")
def _(cov, shots, seed, forced_bits, want_prob):
    B, k = cov.shape[0], cov.shape[-1] // 2
    need_p = want_prob or forced_bits is not None
    return (cov.new_empty((shots, B, k) if forced_bits is None else (0,), dtype=torch.int8),
            cov.new_empty((shots, B) if need_p else (0,), dtype=torch.float64))


@_define("sector_features", "(Tensor cov, Tensor index_sets) -> Tensor")
def _sector_features(cov, index_sets):
    cov = _d(cov, "cov")
    index_sets = _i(index_sets, torch.int32, "index_sets")
    B, D, _ = cov.shape
    T, s = index_sets.shape
    out = torch.empty((B, T), dtype=torch.float64, device=cov.device)
    check(_lib.load().mcb200_sector_features(_p(cov), B, D, _p(index_sets), T, s, _p(out), _stream()))
    return out


@_fake("sector_features")
def _(cov, index_sets):
    return cov.new_empty((cov.shape[0], index_sets.shape[0]), dtype=torch.float64)


@_define("sector_features_backward", "(Tensor cov, Tensor index_sets, Tensor feats, Tensor gf) -> Tensor")
def _sector_features_backward(cov, index_sets, feats, gf):
    cov, feats, gf = _d(cov, "cov"), _d(feats, "feats"), _d(gf, "gf")
    index_sets = _i(index_sets, torch.int32, "index_sets")
    B, D, _ = cov.shape
    T, s = index_sets.shape
    gcov = torch.empty_like(cov)
    check(_lib.load().mcb200_sector_features_backward(_p(cov), B, D, _p(index_sets), T, s, _p(feats), _p(gf), _p(gcov),
                                                      _stream()))
    return gcov


@_fake("sector_features_backward")
def _(cov, index_sets, feats, gf):
    return cov.new_empty(cov.shape, dtype=torch.float64)


def _sf_setup(ctx, inputs, output):
    cov, index_sets = inputs
    ctx.index_sets = index_sets
    ctx.save_for_backward(cov.detach(), output.detach())


def _sf_backward(ctx, gf):
    cov, feats = ctx.saved_tensors
    return torch.ops.mcb200.sector_features_backward(cov, ctx.index_sets, feats, gf.to(torch.float64).contiguous()), None


_tl.register_autograd(f"{NS}::sector_features", _sf_backward, setup_context=_sf_setup)


@_define("rowdot", "(Tensor feats, Tensor coeffs, Tensor? acc) -> Tensor")
def _rowdot(feats, coeffs, acc):
    """out[b] = (acc[b] +) sum_t feats[b, t] coeffs[t]."""
    feats, coeffs = _d(feats, "feats"), _d(coeffs, "coeffs")
    B, T = feats.shape
    out = torch.empty((B,), dtype=torch.float64, device=feats.device) if acc is None else _d(acc, "acc").clone()
    check(_lib.load().mcb200_rowdot(_p(feats), _p(coeffs), B, T, int(acc is not None), _p(out), _stream()))
    return out


@_fake("rowdot")
def _(feats, coeffs, acc):
    return feats.new_empty((feats.shape[0],), dtype=torch.float64)


def _rd_setup(ctx, inputs, output):
    feats, coeffs, acc = inputs
    ctx.has_acc = acc is not None
    ctx.save_for_backward(feats.detach(), coeffs.detach())


def _rd_backward(ctx, g):
    feats, coeffs = ctx.saved_tensors  # linear: outer / matvec products are the adjoint
    return g[:, None] * coeffs[None, :], feats.transpose(0, 1) @ g, (g if ctx.has_acc else None)


_tl.register_autograd(f"{NS}::rowdot", _rd_backward, setup_context=_rd_setup)


@_define("pauli_sum_small", "(Tensor cov, Tensor term_ptr, Tensor indices, Tensor coeffs, Tensor? acc) -> Tensor")
def _pauli_sum_small(cov, term_ptr, indices, coeffs, acc):
    cov, coeffs = _d(cov, "cov"), _d(coeffs, "coeffs")
    term_ptr, indices = _i(term_ptr, torch.int32, "term_ptr"), _i(indices, torch.int32, "indices")
    B, D = cov.shape[0], cov.shape[-1]
    T = coeffs.shape[0]
    if term_ptr.shape[0] != T + 1:
        raise Mcb200Error(f"pauli_sum_small: term_ptr must have {T + 1} entries, got {term_ptr.shape[0]}")
    out = torch.empty((B,), dtype=torch.float64, device=cov.device) if acc is None else _d(acc, "acc").clone()
    check(_lib.load().mcb200_pauli_sum_small(_p(cov), B, D, _p(term_ptr), _p(indices), _p(coeffs), T,
                                             int(acc is not None), _p(out), _stream()))
    return out


@_fake("pauli_sum_small")
def _(cov, term_ptr, indices, coeffs, acc):
    return cov.new_empty((cov.shape[0],), dtype=torch.float64)


# ------------------------------------------------------------------------------------------------ Gram
def _gram_k_pointer(out: torch.Tensor, n0: int, n1: int, rb: int, re: int, rows_only: bool) -> int:
    """Pointer the kernel sees as ``K`` (row 0 of the FULL matrix).  ``rows_only``: ``out`` holds just the rows
    [rb, re) -- e.g. a view into the final matrix of a row-sharded build; the kernel only writes K[i * ldk + j] for
    rb <= i < re (mode TRIANGLE / FULL), so K is the buffer shifted back by rb rows and nothing of size (n0, n1) is
    ever allocated or zero-filled."""
    if out.dtype != torch.float64 or out.stride(1) != 1:
        raise ValueError("Gram buffer must be float64 with unit column stride")
    if rows_only:
        if out.shape != (re - rb, n1):
            raise ValueError(f"rows-only Gram buffer must be ({re - rb}, {n1}), got {tuple(out.shape)}")
        return out.data_ptr() - rb * out.stride(0) * 8 if re > rb else out.data_ptr()
    if out.shape != (n0, n1):
        raise ValueError(f"Gram buffer must be ({n0}, {n1}), got {tuple(out.shape)}")
    return out.data_ptr()


@_define("gram_out", "(Tensor(a!) out, Tensor cov0, Tensor cov1, float scale, int mode, int row_begin, int row_end, "
                     "bool rows_only, float floor2) -> ()")
def _gram_out(out, cov0, cov1, scale, mode, rb, re, rows_only, floor2):
    cov0, cov1 = _d(cov0, "cov0"), _d(cov1, "cov1")
    n0, d, _ = cov0.shape
    n1 = cov1.shape[0]
    if cov1.shape[1:] != (d, d):
        raise ValueError("cov0 and cov1 must hold matrices of the same size")
    kptr = _gram_k_pointer(out, n0, n1, rb, re, rows_only)
    lib = _lib.load()
    ws = _workspace(lib.mcb200_gram_workspace_bytes(n0, n1, d), cov0.device)
    with _Floor(floor2):
        check(lib.mcb200_gram(_p(cov0), _p(cov1), n0, n1, d, rb, re, mode, float(scale), kptr, out.stride(0), _p(ws),
                              _stream()))


@_fake("gram_out")
def _(out, cov0, cov1, scale, mode, rb, re, rows_only, floor2):
    return None


@_define("gram_blocks4_out", "(Tensor(a!) out, Tensor blocks0, Tensor blocks1, int mode, int row_begin, int row_end, "
                             "bool rows_only, float floor2) -> ()")
def _gram_blocks4_out(out, blocks0, blocks1, mode, rb, re, rows_only, floor2):
    blocks0, blocks1 = _d(blocks0, "blocks0"), _d(blocks1, "blocks1")
    n0, Cc, six = blocks0.shape
    n1 = blocks1.shape[0]
    if six != 6 or blocks1.shape[1:] != (Cc, 6):
        raise ValueError("blocks must be (n, C, 6)")
    kptr = _gram_k_pointer(out, n0, n1, rb, re, rows_only)
    with _Floor(floor2):
        check(_lib.load().mcb200_gram_blocks4(_p(blocks0), _p(blocks1), n0, n1, Cc, rb, re, mode, 2.0 ** (-2 * Cc),
                                              kptr, out.stride(0), _stream()))


@_fake("gram_blocks4_out")
def _(out, blocks0, blocks1, mode, rb, re, rows_only, floor2):
    return None


@_define("gram_symmetrize_", "(Tensor(a!) k) -> ()")
def _gram_symmetrize(k):
    if k.dtype != torch.float64 or k.stride(1) != 1:
        raise ValueError("Gram buffer must be float64 with unit column stride")
    check(_lib.load().mcb200_gram_symmetrize(_p(k), k.shape[0], k.shape[1], k.stride(0), _stream()))


@_fake("gram_symmetrize_")
def _(k):
    return None


def signed_seed(seed: int) -> int:
    """A 64-bit seed as the signed integer an ``int`` schema argument can carry."""
    s = int(seed) & _U64
    return s - (1 << 64) if s >= (1 << 63) else s
