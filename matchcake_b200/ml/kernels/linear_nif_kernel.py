"""``LinearNIFKernel`` -- a learnable linear encoder into a free-fermion Hamiltonian, on the GPU.

Mirrors /root/reference/src/matchcake/ml/kernels/linear_nif_kernel.py:14-160: the features go through
``Flatten -> LazyLinear(d(d-1)/2) -> activation``, the outputs fill the strict upper triangle of a real skew-symmetric
``h`` (d = 2N), and the sample's circuit is the single SPTM ``R = exp(h)`` (:107-127).  The Gram matrix is then the one
of :class:`NIFKernel` -- per-sample covariances ``R^T Lambda_0 R`` and ``K_ij = 2^-N |Pf(Lambda_i + Lambda_j)|``.

The reference calls ``torch.matrix_exp``; here the exponential is a scaling-and-squaring Taylor series whose every matrix
product is the library's batched FP64 DMMA chain step (``mcb200_sptm_matmul`` through ``torch.ops.mcb200.sptm_matmul``,
which also carries the adjoint), so the encoder is trainable through kernel alignment without leaving the CUDA path.
"""
from __future__ import annotations

import math
from typing import Optional

import numpy as np
import torch

from ... import ops as _ops
from ...operations import SingleParticleTransitionMatrixOperation
from .nif_kernel import NIFKernel

_TAYLOR_DEGREE = 16          # remainder (1/2)^17 / 17! = 2e-20 for ||a||_1 <= 1/2
_INV_FACT = [1.0 / math.factorial(k) for k in range(_TAYLOR_DEGREE + 1)]


def skew_expm(h: torch.Tensor) -> torch.Tensor:
    """exp(h) for a batch of real (skew-symmetric) matrices (B, d, d), float64 on the GPU.

    ``a = h / 2^s`` with ``||a||_1 <= 1/2``; degree-16 Taylor polynomial by Paterson-Stockmeyer in ``a^4`` (7 products),
    then ``s`` squarings -- every product is ``ops.sptm_matmul``."""
    if h.ndim == 2:
        return skew_expm(h[None])[0]
    h = h.to(torch.float64)
    norm = float(h.detach().abs().sum(dim=-2).amax()) if h.numel() else 0.0
    s = max(0, int(math.ceil(math.log2(norm / 0.5)))) if norm > 0.5 else 0
    a = h * (2.0 ** -s)
    mm = _ops.sptm_matmul
    eye = torch.eye(h.shape[-1], dtype=torch.float64, device=h.device).expand_as(a)
    a2 = mm(a, a)
    a3 = mm(a2, a)
    a4 = mm(a2, a2)
    c = _INV_FACT

    def q(j):  # sum_{i<4} c[4j+i] a^i
        out = c[4 * j] * eye
        if 4 * j + 1 <= _TAYLOR_DEGREE:
            out = out + c[4 * j + 1] * a + c[4 * j + 2] * a2 + c[4 * j + 3] * a3
        return out

    p = q(4)
    for j in (3, 2, 1, 0):
        p = mm(p, a4) + q(j)
    for _ in range(s):
        p = mm(p, p)
    return p


class LinearNIFKernel(NIFKernel):
    DEFAULT_BIAS = True
    DEFAULT_ENCODER_ACTIVATION = "Identity"

    def __init__(self, *, gram_batch_size: int = NIFKernel.DEFAULT_GRAM_BATCH_SIZE,
                 random_state: int = NIFKernel.DEFAULT_RANDOM_STATE, alignment: bool = NIFKernel.DEFAULT_ALIGNMENT,
                 alignment_iterations: int = NIFKernel.DEFAULT_ALIGNMENT_ITERATIONS,
                 alignment_learning_rate: float = NIFKernel.DEFAULT_ALIGNMENT_LEARNING_RATE,
                 alignment_early_stopping_patience: int = NIFKernel.DEFAULT_ALIGNMENT_EARLY_STOPPING_PATIENCE,
                 alignment_early_stopping_threshold: float = NIFKernel.DEFAULT_ALIGNMENT_EARLY_STOPPING_THRESHOLD,
                 n_qubits: int = NIFKernel.DEFAULT_N_QUBITS, r_dtype: Optional[torch.dtype] = None,
                 c_dtype: Optional[torch.dtype] = None, bias: bool = DEFAULT_BIAS,
                 encoder_activation: str = DEFAULT_ENCODER_ACTIVATION):
        super().__init__(gram_batch_size=gram_batch_size, random_state=random_state, alignment=alignment,
                         alignment_iterations=alignment_iterations, alignment_learning_rate=alignment_learning_rate,
                         alignment_early_stopping_patience=alignment_early_stopping_patience,
                         alignment_early_stopping_threshold=alignment_early_stopping_threshold, n_qubits=n_qubits,
                         r_dtype=r_dtype, c_dtype=c_dtype)
        self._bias = bool(bias)
        self._encoder_activation = encoder_activation
        self.encoder = torch.nn.Sequential(torch.nn.Flatten(), self._new_linear(),
                                           getattr(torch.nn, self._encoder_activation)())

    def _new_linear(self) -> torch.nn.Module:
        return torch.nn.LazyLinear(self.encoder_out_indices[0].size, bias=self._bias, dtype=self.R_DTYPE)

    # ------------------------------------------------------------------ sklearn-visible properties (:129-160)
    @property
    def n_qubits(self) -> int:
        return self._n_qubits

    @n_qubits.setter
    def n_qubits(self, value: int):
        self._n_qubits = int(value)
        self._q_device = None
        if "encoder" in self._modules:
            self.encoder[1] = self._new_linear()

    @property
    def bias(self) -> bool:
        return self._bias

    @bias.setter
    def bias(self, value: bool):
        self._bias = bool(value)
        self.encoder[1] = self._new_linear()

    @property
    def encoder_activation(self) -> str:
        return self._encoder_activation

    @encoder_activation.setter
    def encoder_activation(self, value: str):
        self._encoder_activation = value
        self.encoder[-1] = getattr(torch.nn, self._encoder_activation)()

    @property
    def encoder_out_indices(self):
        return np.triu_indices(2 * self._n_qubits, k=1)

    @property
    def encoder_out_tril_indices(self):
        return np.tril_indices(2 * self._n_qubits, k=-1)

    # ------------------------------------------------------------------ ansatz
    def hamiltonian(self, x) -> torch.Tensor:
        """(n_samples, d, d) real skew-symmetric generator: encoder outputs above the diagonal, minus them below."""
        self.encoder.to("cuda")
        x = torch.as_tensor(np.asarray(x) if not isinstance(x, torch.Tensor) else x).to(device="cuda",
                                                                                        dtype=self.R_DTYPE)
        d = 2 * self._n_qubits
        iu = self.encoder_out_indices
        upper = torch.zeros((x.shape[0], d, d), dtype=torch.float64, device="cuda")
        upper[:, torch.as_tensor(iu[0], device="cuda"), torch.as_tensor(iu[1], device="cuda")] = \
            self.encoder(x).to(torch.float64)
        return upper - upper.transpose(1, 2)

    def _x_to_sptm(self, x) -> torch.Tensor:
        # the encoder's parameters always require grad; only kernel alignment (train mode) differentiates through them
        with torch.set_grad_enabled(self._differentiating()):
            return skew_expm(self.hamiltonian(x))

    def ansatz(self, x):
        """One SPTM operation per sample batch (linear_nif_kernel.py:107-127)."""
        yield SingleParticleTransitionMatrixOperation(self._x_to_sptm(x), wires=self.wires)
