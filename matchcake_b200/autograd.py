"""Reverse-mode glue: ``torch.autograd.Function`` wrappers whose backward passes are the hand-written adjoint kernels of
``csrc/backward.cu`` (include/mcb200.h, "Reverse-mode derivatives").

The reference differentiates this path with torch autograd over its op sequence (``devices/nif_device.py:949-989``
advertises backprop, ``ml/kernels/kernel.py:200-283`` trains ``bias_`` through the Gram matrix, torch_pfaffian supplies
``d Pf = Pf/2 tr(A^-1 dA)``).  Here the forward kernels are opaque to autograd, so every differentiable op of
``matchcake_b200.ops`` routes through one of these functions when (and only when) an input requires grad; the
inference path never pays for it.  Nothing is recomputed on the host and there is no torch fallback: a missing adjoint
raises ``NotImplementedError``.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import check


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def needs_grad(*tensors) -> bool:
    return torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad for t in tensors)


class CircuitColumns(torch.autograd.Function):
    """X = R[:, cols]; backward un-applies the gates (mcb200_circuit_columns_backward)."""

    @staticmethod
    def forward(ctx, params, circ, cols):
        from . import ops
        x = ops.circuit_columns(circ, params.detach(), cols)
        ctx.circ = circ
        ctx.save_for_backward(params.detach(), x)
        return x

    @staticmethod
    def backward(ctx, gx):
        params, x = ctx.saved_tensors
        circ = ctx.circ
        gx = gx.to(torch.float64).contiguous()
        B, _, m = x.shape
        ga = torch.empty((B, circ.n_param_cols), dtype=torch.float64, device=x.device)
        check(_lib.load().mcb200_circuit_columns_backward(C.byref(circ._struct), params.data_ptr(), B, m, x.data_ptr(),
                                                          gx.data_ptr(), ga.data_ptr(), _stream()))
        if circ.scale_dev is not None:  # angle = bias + scale * x was applied in-kernel
            ga = ga * circ.scale_dev[None, :]
        return ga, None, None


class CovBasis(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, n_qubits, bits):
        from . import ops
        xd = x.detach()
        cov = ops.cov_basis(xd, n_qubits, bits)
        ctx.n_qubits, ctx.bits = n_qubits, bits
        ctx.save_for_backward(xd.contiguous())
        return cov

    @staticmethod
    def backward(ctx, gcov):
        (x,) = ctx.saved_tensors
        gcov = gcov.to(torch.float64).contiguous()
        B, _, m = x.shape
        bits = ctx.bits
        ib = None if bits is None else bits.to(device=x.device, dtype=torch.int8).contiguous()
        gx = torch.empty_like(x)
        check(_lib.load().mcb200_cov_basis_backward(x.data_ptr(), None if ib is None else ib.data_ptr(), B,
                                                    ctx.n_qubits, m, gcov.data_ptr(), gx.data_ptr(), _stream()))
        return gx, None, None


class CovDense(torch.autograd.Function):
    """cov = (R~^T L0 R~)[idx, idx] for product-state / extended covariances; adjoint w.r.t. R only (the initial-state
    covariance is treated as a constant)."""

    @staticmethod
    def forward(ctx, r, l0, idx):
        from . import ops
        rd, ld = r.detach(), l0.detach()
        ctx.idx = idx
        ctx.r_shape = tuple(r.shape)
        ctx.save_for_backward(rd, ld)
        return ops.cov_dense(rd, ld, idx)

    @staticmethod
    def backward(ctx, gcov):
        from . import ops
        r, l0 = ctx.saved_tensors
        r3 = r if r.ndim == 3 else r[None]
        r3 = ops._f64(r3, "r")
        l0 = ops._f64(l0, "l0")
        Br, d, _ = r3.shape
        D = l0.shape[-1]
        gcov = gcov.to(torch.float64).contiguous()
        B, m, _ = gcov.shape
        sr = d * d if (Br == B and B > 1) else 0
        sl = D * D if (l0.ndim == 3 and l0.shape[0] == B and B > 1) else 0
        idx = None if ctx.idx is None else ctx.idx.to(device=r3.device, dtype=torch.int32).contiguous()
        lib = _lib.load()
        ws = torch.empty((lib.mcb200_cov_dense_backward_workspace_bytes(B, D, m) + 7) // 8, dtype=torch.float64,
                         device=r3.device)
        gr = torch.empty((B if sr else 1, d, d), dtype=torch.float64, device=r3.device)
        check(lib.mcb200_cov_dense_backward(r3.data_ptr(), sr, l0.data_ptr(), sl, B, d, D,
                                            None if idx is None else idx.data_ptr(), m, gcov.data_ptr(), gr.data_ptr(),
                                            ws.data_ptr(), _stream()))
        # a shared R (2-D, or batch of one) receives the sum over the batch: the kernel accumulates it when strideR == 0
        return gr.reshape(ctx.r_shape), None, None


class Pfaffian(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, signed, scale):
        from . import ops
        ad = a.detach().to(torch.float64).contiguous()
        out = ops.pfaffian(ad, signed, scale)
        ctx.save_for_backward(ad, out)
        return out

    @staticmethod
    def backward(ctx, gout):
        a, out = ctx.saved_tensors
        m = a.shape[-1]
        n = out.numel()
        ga = torch.empty_like(a)
        if n and m:
            g = gout.to(torch.float64).contiguous()
            check(_lib.load().mcb200_pfaffian_backward(a.data_ptr(), n, m, out.contiguous().data_ptr(), g.data_ptr(),
                                                       ga.data_ptr(), _stream()))
        return ga, None, None


class Probs(torch.autograd.Function):
    """Unnormalised outcome probabilities; the row normalisation is done by the caller with differentiable torch ops."""

    @staticmethod
    def forward(ctx, cov, outcomes):
        from . import ops
        cd = cov.detach().to(torch.float64).contiguous()
        p = ops.probs(cd, outcomes, normalize=False)
        ctx.outcomes = outcomes.to(device=cd.device, dtype=torch.int8).contiguous()
        ctx.save_for_backward(cd, p)
        return p

    @staticmethod
    def backward(ctx, gp):
        cov, p = ctx.saved_tensors
        B, Q = p.shape
        k = cov.shape[-1] // 2
        g = gp.to(torch.float64).contiguous()
        gcov = torch.empty_like(cov)
        check(_lib.load().mcb200_probs_backward(cov.data_ptr(), ctx.outcomes.data_ptr(), B, Q, k, p.data_ptr(),
                                                g.data_ptr(), gcov.data_ptr(), _stream()))
        return gcov, None


class SectorFeatures(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cov, index_sets):
        from . import ops
        cd = cov.detach().to(torch.float64).contiguous()
        feats = ops.sector_features(cd, index_sets)
        ctx.index_sets = index_sets.to(device=cd.device, dtype=torch.int32).contiguous()
        ctx.save_for_backward(cd, feats)
        return feats

    @staticmethod
    def backward(ctx, gf):
        cov, feats = ctx.saved_tensors
        B, D, _ = cov.shape
        T, s = ctx.index_sets.shape
        g = gf.to(torch.float64).contiguous()
        gcov = torch.empty_like(cov)
        check(_lib.load().mcb200_sector_features_backward(cov.data_ptr(), B, D, ctx.index_sets.data_ptr(), T, s,
                                                          feats.data_ptr(), g.data_ptr(), gcov.data_ptr(), _stream()))
        return gcov, None


class SptmMatmul(torch.autograd.Function):
    """C = op(a) op(b); the adjoint is two more chain steps on the same DMMA kernel."""

    @staticmethod
    def forward(ctx, a, b, trans_a, trans_b):
        from . import ops
        ad, bd = a.detach(), b.detach()
        ctx.trans = (trans_a, trans_b)
        ctx.save_for_backward(ad, bd)
        return ops.sptm_matmul(ad, bd, trans_a, trans_b)

    @staticmethod
    def backward(ctx, gc):
        from . import ops
        a, b = ctx.saved_tensors
        ta, tb = ctx.trans
        gc = gc.to(torch.float64).contiguous()
        if gc.ndim == 2:
            gc = gc[None]
        # C = A' B' with A' = op(a), B' = op(b):  gA' = gC B'^T,  gB' = A'^T gC
        ga = ops.sptm_matmul(gc, b, False, not tb) if not ta else ops.sptm_matmul(b, gc, tb, True)
        gb = ops.sptm_matmul(a, gc, not ta, False) if not tb else ops.sptm_matmul(gc, a, True, ta)
        if a.ndim == 2:
            ga = ga.sum(0) if ga.ndim == 3 else ga
        if b.ndim == 2:
            gb = gb.sum(0) if gb.ndim == 3 else gb
        return ga, gb, None, None
