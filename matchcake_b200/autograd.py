"""Reverse mode of the CUDA path.

The reference differentiates this path with torch autograd over its op sequence (``devices/nif_device.py:949-989``
advertises backprop, ``ml/kernels/kernel.py:200-283`` trains ``bias_`` through the Gram matrix, torch_pfaffian supplies
``d Pf = Pf/2 tr(A^-1 dA)``).  Here the forward kernels are opaque to autograd, so every differentiable operator of
``torch.ops.mcb200`` carries a ``torch.library.register_autograd`` formula (``matchcake_b200/torch_ops.py``) whose
backward pass is one of the hand-written adjoint kernels of ``csrc/backward.cu`` (include/mcb200.h, "Reverse-mode
derivatives"), itself registered as a ``mcb200::*_backward`` operator.  Nothing is recomputed on the host and there is
no torch fallback: a missing adjoint raises ``NotImplementedError``.

The fused inference kernels (projector expval, outcome tree, small-sector Pauli sums, in-kernel magnitude floor) have no
adjoint; ``matchcake_b200.ops`` asks :func:`needs_grad` and routes through the unfused differentiable operators when --
and only when -- an input requires grad, so inference never pays for it.
"""
from __future__ import annotations

import torch


def needs_grad(*tensors) -> bool:
    return torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad for t in tensors)
