"""``NIFKernel`` -- fermionic kernel Gram builder on the GPU.

Mirrors /root/reference/src/matchcake/ml/kernels/nif_kernel.py:17-264.  The reference evaluates every Gram cell
as its own circuit (``R = sptm0[i] @ sptm1[j]^T``, complex lookup-table observable, one LU per pair) driven by a
Python loop over index pairs.  Here:

  1. the ansatz is compiled ONCE into a gate-list circuit and the per-sample SPTMs ``R_i`` come from one kernel
     launch (``_x_to_sptm``), then the per-sample covariances ``Lambda_i = R_i^T Lambda_0 R_i`` from another;
  2. every cell is ``K_ij = 2^-N |Pf(Lambda_i + Lambda_j)|`` (identical to
     ``P(0..0 | R_i R_j^T)``), evaluated by the Gram kernel with the reference's triangle / mirror / unit
     diagonal rules applied in-kernel;
  3. with ``torch.distributed`` initialised the rows are sharded over the ranks and assembled by one all-gather.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch

from ... import distributed as dist_mod
from ... import ops as _ops
from ...device import NonInteractingFermionicDevice
from ...operations import SingleParticleTransitionMatrixOperation
from .gram_matrix import GramMatrix  # noqa: F401  (re-exported; semantics implemented in-kernel)
from .kernel import Kernel


class NIFKernel(Kernel):
    DEFAULT_N_QUBITS = 12
    DEFAULT_R_DTYPE = torch.float32
    DEFAULT_C_DTYPE = torch.complex64

    def __init__(self, *, gram_batch_size: int = Kernel.DEFAULT_GRAM_BATCH_SIZE,
                 random_state: int = Kernel.DEFAULT_RANDOM_STATE, alignment: bool = Kernel.DEFAULT_ALIGNMENT,
                 alignment_iterations: int = Kernel.DEFAULT_ALIGNMENT_ITERATIONS,
                 alignment_learning_rate: float = Kernel.DEFAULT_ALIGNMENT_LEARNING_RATE,
                 alignment_early_stopping_patience: int = Kernel.DEFAULT_ALIGNMENT_EARLY_STOPPING_PATIENCE,
                 alignment_early_stopping_threshold: float = Kernel.DEFAULT_ALIGNMENT_EARLY_STOPPING_THRESHOLD,
                 n_qubits: int = DEFAULT_N_QUBITS, r_dtype: Optional[torch.dtype] = None,
                 c_dtype: Optional[torch.dtype] = None):
        super().__init__(gram_batch_size=gram_batch_size, random_state=random_state, alignment=alignment,
                         alignment_iterations=alignment_iterations, alignment_learning_rate=alignment_learning_rate,
                         alignment_early_stopping_patience=alignment_early_stopping_patience,
                         alignment_early_stopping_threshold=alignment_early_stopping_threshold)
        self._r_dtype = r_dtype or self.DEFAULT_R_DTYPE
        self._c_dtype = c_dtype or self.DEFAULT_C_DTYPE
        self._n_qubits = int(n_qubits)
        self._q_device = None
        self._covs_train = None
        #: store the Gram in float32 like the reference's GramMatrix (gram_matrix.py:38-47); False keeps float64
        self.float32_gram = True
        #: when the ansatz never couples different wire pairs the covariance is block diagonal (4x4 blocks) and the
        #: Gram factorises into 2-qubit factors (csrc/gram_blocks.cu); False forces the dense Pfaffian path
        self.use_block_factorisation = True
        self._blocks_train = None
        #: accuracy / speed of the d <= 64 Gram kernels (ops.gram): "auto" evaluates a sample of the cells with the fast and
        #: the strict pivot bound and keeps the fast one only if they agree to 1e-11; "fast" / "strict" force one
        self.gram_pivoting = "auto"
        #: optional list that receives (stage, cuda event) pairs of the multi-GPU Gram build (bench / profiling)
        self.timeline = None
        #: several ranks: every rank evolves only its slice of the samples and the covariances are all-gathered;
        #: False = every rank evolves all samples (no collective before the Gram kernel)
        self.shard_precompute = True
        #: EPSILON of the reference's magnitude Pfaffian: its Gram entries are sqrt(K^2 + 1e-32) (each cell is a
        #: LookupTableStrategy probability, nif_kernel.py:210-221 -> lookup_table_strategy.py:70-71); 0 = exact |Pf|
        self.pfaffian_epsilon = 1e-32

    # sklearn's get_params reads attributes named like the __init__ arguments
    @property
    def n_qubits(self) -> int:
        return self._n_qubits

    @n_qubits.setter
    def n_qubits(self, value: int):
        self._n_qubits = int(value)
        self._q_device = None

    @property
    def r_dtype(self):
        return self._r_dtype

    @property
    def c_dtype(self):
        return self._c_dtype

    @property
    def R_DTYPE(self):
        return self._r_dtype

    @property
    def C_DTYPE(self):
        return self._c_dtype

    @property
    def q_device(self) -> NonInteractingFermionicDevice:
        if self._q_device is None:
            self._q_device = NonInteractingFermionicDevice(self._n_qubits, r_dtype=self._r_dtype,
                                                           c_dtype=self._c_dtype, output_device="cuda")
        return self._q_device

    @property
    def wires(self):
        return tuple(range(self._n_qubits))

    # ------------------------------------------------------------------ reference API
    def forward(self, x0, x1=None, **kwargs) -> torch.Tensor:
        if x1 is None:
            x1 = x0
        return self.compute_similarities(x0, x1, **kwargs)

    def ansatz(self, x):
        raise NotImplementedError

    def circuit(self, sptm0, sptm1):
        """The per-pair circuit of the reference (nif_kernel.py:159-177), kept for API compatibility."""
        yield SingleParticleTransitionMatrixOperation(sptm0, wires=self.wires)
        yield SingleParticleTransitionMatrixOperation(sptm1, wires=self.wires).adjoint()

    def _x_to_sptm(self, x) -> torch.Tensor:
        """(n_samples, 2N, 2N) SPTMs; generic route through the device for subclasses that only define
        ``ansatz`` (FermionicPQCKernel overrides this with a compiled circuit)."""
        dev = self.q_device
        dev.execute_generator(self.ansatz(x), reset=True)
        r = dev._materialize()
        return r.expand(len(x), -1, -1) if r.shape[0] == 1 else r

    def _x_to_cov(self, x) -> torch.Tensor:
        """Per-sample covariance Lambda_i = R_i^T Lambda_0 R_i for the |0..0> input (n_samples, 2N, 2N).  With several
        ranks every rank evolves its slice of the samples and one all-gather replicates the covariances (the Gram
        cells of a rank need all of them)."""
        rank, ws = dist_mod.world()
        n = len(x)
        if ws == 1 or n < 4 * ws or self._differentiating() or not self.shard_precompute:
            return _ops.cov_basis(self._x_to_sptm(x), self._n_qubits)
        ranges = dist_mod.even_ranges(n, ws)
        s, e = ranges[rank]
        d = 2 * self._n_qubits
        cov = torch.empty((n, d, d), dtype=torch.float64, device="cuda")
        _ops.cov_basis(self._x_to_sptm(x[s:e]), self._n_qubits, out=cov[s:e])   # straight into the final rows
        if self.timeline is not None:
            self.timeline.append(("cov_local", dist_mod._event()))
        if n % ws == 0:
            import torch.distributed as dist

            dist.all_gather_into_tensor(cov, cov[s:e])                          # in place: slot `rank` is the input
        else:
            dist_mod.gather_row_ranges(cov, ranges)
        if self.timeline is not None:
            self.timeline.append(("cov_all_gather", dist_mod._event()))
        return cov

    @property
    def sptms_train(self) -> torch.Tensor:
        return self._x_to_sptm(self.x_train_)

    def _state_key(self):
        """Identity of the training inputs and the version of every parameter: cached per-sample data is stale once an
        optimiser step (kernel alignment) has touched a parameter."""
        return (id(self.x_train_),) + tuple((p.data_ptr(), p._version) for p in self.parameters())

    @property
    def covs_train(self) -> torch.Tensor:
        key = self._state_key()
        if self._covs_train is None or self._covs_train[0] != key:
            self._covs_train = (key, self._x_to_cov(self.x_train_))
        return self._covs_train[1]

    def _differentiating(self) -> bool:
        """Train mode with a parameter that requires grad: take the differentiable route."""
        return self.training and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())

    def _block_plan(self):
        """Sub-circuits per independent wire pair, or None (subclasses with a compiled ansatz override this)."""
        return None

    def _x_to_blocks(self, x) -> torch.Tensor:
        raise NotImplementedError

    def compute_similarities(self, x0, x1, **kwargs) -> torch.Tensor:
        same = x1 is x0
        full = kwargs.get("full_rectangle", False)
        if self._differentiating():
            # kernel alignment (kernel.py:200-283): K1 -> K2 -> Pfaffian through the ops that have adjoint kernels
            cov0 = self._x_to_cov(x0)
            cov1 = cov0 if same else self._x_to_cov(x1)
            k = _ops.gram(cov0, cov1, 2.0 ** (-self._n_qubits), mode=_ops.GRAM_FULL if full else _ops.GRAM_REFERENCE,
                          floor2=self.pfaffian_epsilon)
            return k.to(self.R_DTYPE)
        if self.use_block_factorisation and self._block_plan() is not None:
            b0 = self._x_to_blocks(x0)
            if kwargs.get("x1_train", False):
                key = self._state_key()
                if self._blocks_train is None or self._blocks_train[0] != key:
                    self._blocks_train = (key, self._x_to_blocks(self.x_train_))
                b1 = self._blocks_train[1]
            else:
                b1 = b0 if same else self._x_to_blocks(x1)
            k = self.gram_from_blocks(b0, b1, full=full, gather=kwargs.get("gather", True))
            if isinstance(k, tuple):  # row-sharded: (local rows, (begin, end))
                return (k[0].to(torch.float32) if self.float32_gram else k[0]).to(self.R_DTYPE), k[1]
            if self.float32_gram:
                k = k.to(torch.float32)
            return k.to(self.R_DTYPE)
        cov0 = self._x_to_cov(x0)
        if kwargs.get("x1_train", False):
            cov1 = self.covs_train
        else:
            cov1 = cov0 if same else self._x_to_cov(x1)
        k = self.gram_from_covariances(cov0, cov1, full=full, gather=kwargs.get("gather", True))
        if isinstance(k, tuple):
            return (k[0].to(torch.float32) if self.float32_gram else k[0]).to(self.R_DTYPE), k[1]
        if self.float32_gram:
            k = k.to(torch.float32)
        return k.to(self.R_DTYPE)

    def gram_from_covariances(self, cov0: torch.Tensor, cov1: torch.Tensor, full: bool = False,
                              gather: bool = True) -> torch.Tensor:
        n0, n1 = cov0.shape[0], cov1.shape[0]
        scale = 2.0 ** (-self._n_qubits)
        rank, ws = dist_mod.world()
        if ws == 1:
            return _ops.gram(cov0, cov1, scale, mode=_ops.GRAM_FULL if full else _ops.GRAM_REFERENCE,
                             floor2=self.pfaffian_epsilon, pivoting=self.gram_pivoting)

        # one decision per build (not per row block): every rank and every launch takes the same route
        growth = _ops._gram_growth_for(cov0, cov1, n0 * n1, self.gram_pivoting)

        def rows(s, e, out_rows):
            _ops.gram(cov0, cov1, scale, mode=_ops.GRAM_FULL if full else _ops.GRAM_TRIANGLE, row_range=(s, e),
                      out=out_rows, rows_only=True, floor2=self.pfaffian_epsilon, pivoting=growth)

        return self._sharded_gram(rows, n0, n1, full, cov0.device, gather)

    def gram_from_blocks(self, b0: torch.Tensor, b1: torch.Tensor, full: bool = False,
                         gather: bool = True) -> torch.Tensor:
        """Same contract as :meth:`gram_from_covariances` for block-diagonal covariances given as (n, C, 6)."""
        n0, n1 = b0.shape[0], b1.shape[0]
        rank, ws = dist_mod.world()
        if ws == 1:
            return _ops.gram_blocks4(b0, b1, mode=_ops.GRAM_FULL if full else _ops.GRAM_REFERENCE,
                                     floor2=self.pfaffian_epsilon)

        def rows(s, e, out_rows):
            _ops.gram_blocks4(b0, b1, mode=_ops.GRAM_FULL if full else _ops.GRAM_TRIANGLE, row_range=(s, e),
                              out=out_rows, rows_only=True, floor2=self.pfaffian_epsilon)

        return self._sharded_gram(rows, n0, n1, full, b0.device, gather)

    def _sharded_gram(self, rows, n0: int, n1: int, full: bool, device, gather: bool = True):
        """Row-sharded Gram build.  ``rows(begin, end, out_rows)`` launches the Gram kernel for rows
        [begin, end) writing straight into ``out_rows`` = those rows of the final matrix: no (n0, n1) staging buffer,
        no zero fill, no pad / permute copies -- the only traffic besides the kernel's own stores is the all-gather
        and the mirror pass.  ``gather=False`` keeps the result row-sharded and returns ``(local_rows, (begin, end))``
        (Nystroem.transform applies the normalisation per rank and gathers only on request)."""
        rank, ws = dist_mod.world()
        tl = self.timeline
        if n0 == n1 and not full and gather:
            if n0 % (2 * ws) == 0:
                k = dist_mod.folded_rows_in_place(rows, n0, n1, torch.float64, device, timeline=tl)
            else:
                def rows_alloc(s, e):
                    out = torch.empty((e - s, n1), dtype=torch.float64, device=device)
                    rows(s, e, out)
                    return out

                k = dist_mod.sharded_folded_rows(rows_alloc, n0)
            k = _ops.gram_symmetrize(k)
            if tl is not None:
                tl.append(("symmetrize", dist_mod._event()))
            return k
        ranges = dist_mod.balanced_row_ranges(n0, n1, ws, full=full)
        s, e = ranges[rank]
        q = min(n0, n1)
        # the mirror / unit-diagonal pass touches the leading q x q block only; row-sharded output is possible when
        # that block lies inside ONE rank's rows (always true for the Nystroem transform, n0 >> n1)
        if not gather and (full or (n0 >= n1 and ranges[0][1] >= q)):
            local = torch.empty((e - s, n1), dtype=torch.float64, device=device)
            rows(s, e, local)
            if not full and s == 0 and e > 0:
                _ops.gram_symmetrize(local[:q, :q])
            return local, (s, e)
        k = torch.empty((n0, n1), dtype=torch.float64, device=device)
        rows(s, e, k[s:e])
        dist_mod.gather_row_ranges(k, ranges)
        k = k if full else _ops.gram_symmetrize(k)
        return k if gather else (k[s:e], (s, e))

    def _compute_similarities_func(self, indices, sptm0, sptm1, p_bar=None):
        """Expectation values for explicit index pairs (nif_kernel.py:210-221), float64, on the GPU."""
        idx = torch.as_tensor(np.asarray(indices), device=sptm0.device)
        c0 = _ops.cov_basis(sptm0[idx[:, 0]].contiguous(), self._n_qubits)
        c1 = _ops.cov_basis(sptm1[idx[:, 1]].contiguous(), self._n_qubits)
        scale = 2.0 ** (-self._n_qubits)
        return _ops.pfaffian(c0 + c1, signed=False, scale=scale, epsilon=self.pfaffian_epsilon / (scale * scale))
