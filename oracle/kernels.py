"""sklearn-style fermionic kernels: Gram-matrix builder and Nystroem.  TEST INFRASTRUCTURE.

Restates (relative to /root/reference/src/matchcake/ml/kernels):
* ``FermionicPQCKernel.fit`` / ``ansatz`` ........ fermionic_pqc_kernel.py:141-199
* ``SptmAngleEmbedding.compute_decomposition`` ... ../../operations/single_particle_transition_matrices/sptm_angle_embedding.py:17-45
* ``NIFKernel._x_to_sptm`` / ``circuit`` / ``_compute_similarities_func`` / ``compute_similarities`` .. nif_kernel.py:124-221
* ``GramMatrix`` (index generator, float32 storage, symmetrize, unit diagonal) .. gram_matrix.py:12-209
* ``Nystroem.fit`` / ``transform`` ............... nystroem.py:58-123
"""

from __future__ import annotations

from typing import Iterator, List, Optional, Tuple

import numpy as np

from . import gates
from .gates import Op
from .readout import probs_lookup_table

ROT = {"X": "SptmCompRxRx", "Y": "SptmCompRyRy", "Z": "SptmCompRzRz"}


def pattern_double(n):  # fermionic_pqc_kernel.py:16-19
    return [(i, i + 1) for i in range(0, n - 1, 2)]


def pattern_double_odd(n):
    return [(i, i + 1) for i in range(1, n - 1, 2)]


def angle_embedding_ops(sub_x: np.ndarray, n_wires: int, rotations: str) -> List[Op]:
    """sptm_angle_embedding.py:17-45,47-70: pad to even, wires[:n_features], per pair i and per rotation."""
    sub_x = np.asarray(sub_x, dtype=np.float64)
    if sub_x.shape[-1] % 2 != 0:
        sub_x = np.concatenate([sub_x, np.zeros_like(sub_x[..., :1])], axis=-1)
    n_features = sub_x.shape[-1]
    if n_features > n_wires:
        raise ValueError(f"Features must be of length {n_wires} or less; got length {n_features}.")
    ops = []
    for i in range(n_features // 2):
        p = np.stack([sub_x[..., 2 * i], sub_x[..., 2 * i + 1]], axis=-1)
        for r in rotations.split(","):
            ops.append(Op(ROT[r], (2 * i, 2 * i + 1), params=p))
    return ops


class FermionicPQCKernelOracle:
    """Reference-faithful CPU kernel (float64 throughout, i.e. ``r_dtype=float64, c_dtype=complex128``)."""

    def __init__(self, n_qubits: int, rotations: str = "Y,Z", entangling_mth: str = "fswap",
                 random_state: int = 0, gram_batch_size: int = 10_000):
        self.n_qubits = n_qubits
        self.rotations = rotations
        self.entangling_mth = entangling_mth
        self.gram_batch_size = gram_batch_size
        self.np_rn_gen = np.random.RandomState(seed=random_state)  # kernel.py:71
        self.depth_: Optional[int] = None
        self.bias_ = None
        self.data_scaling_ = None
        self.x_train_ = None

    def fit(self, x_train, y_train=None):
        """fermionic_pqc_kernel.py:141-164."""
        x_arr = np.asarray(x_train)
        n_inputs = int(np.prod(x_arr.shape[1:]))
        self.bias_ = self.np_rn_gen.random(n_inputs)
        self.data_scaling_ = np.pi * np.ones(n_inputs)
        if self.depth_ is None:
            self.depth_ = int(max(1, np.ceil(x_arr.shape[-1] / self.n_qubits)))
        self.x_train_ = x_arr
        return self

    def ansatz(self, x: np.ndarray) -> List[Op]:
        """fermionic_pqc_kernel.py:166-199."""
        x = np.asarray(x, dtype=np.float64).reshape(len(x), -1)
        x = self.bias_ + self.data_scaling_ * x
        patterns = [pattern_double(self.n_qubits), pattern_double_odd(self.n_qubits)]
        ops: List[Op] = []
        for layer in range(self.depth_):
            sub_x = x[..., layer * self.n_qubits:(layer + 1) * self.n_qubits]
            ops.extend(angle_embedding_ops(sub_x, self.n_qubits, self.rotations))
            for wires in patterns[layer % 2]:
                if self.entangling_mth == "fswap":
                    ops.append(Op("SptmFSwap", wires))
                elif self.entangling_mth == "identity":
                    pass
                else:
                    raise ValueError(f"Unknown entangling method: {self.entangling_mth}")
        return ops

    def x_to_sptm(self, x: np.ndarray) -> np.ndarray:
        """nif_kernel.py:179-208 (chunks of gram_batch_size)."""
        x = np.asarray(x, dtype=np.float64)
        chunks = np.array_split(np.arange(len(x)), int(np.ceil(len(x) / self.gram_batch_size)))
        d = 2 * self.n_qubits
        out = [np.broadcast_to(gates.chain(self.ansatz(x[c]), self.n_qubits), (len(c), d, d)) for c in chunks]
        return np.concatenate(out, axis=0)

    def pair_expvals(self, sptm0: np.ndarray, sptm1: np.ndarray, indices: np.ndarray) -> np.ndarray:
        """``_compute_similarities_func`` (nif_kernel.py:210-221): R = b0 @ b1^T, then <0..0|rho|0..0>
        through the lookup-table strategy.  float64 values (before the float32 Gram store)."""
        b0, b1 = sptm0[indices[:, 0]], sptm1[indices[:, 1]]
        r = np.einsum("...ij,...jk->...ik", b0, np.swapaxes(b1, -1, -2))
        zeros = np.zeros(self.n_qubits, dtype=int)
        return probs_lookup_table(r, zeros, zeros, np.arange(self.n_qubits))

    def gram(self, x0: np.ndarray, x1: Optional[np.ndarray] = None, float32_store: bool = True) -> np.ndarray:
        """``compute_similarities`` (nif_kernel.py:124-157)."""
        x1 = x0 if x1 is None else x1
        sptm0 = self.x_to_sptm(x0)
        sptm1 = sptm0 if x1 is x0 else self.x_to_sptm(x1)
        g = GramMatrixOracle((len(x0), len(x1)), dtype=np.float32 if float32_store else np.float64)
        g.apply_(lambda idx: self.pair_expvals(sptm0, sptm1, idx), batch_size=self.gram_batch_size)
        return g.data

    def transform(self, x: np.ndarray, **kw) -> np.ndarray:
        return self.gram(x, self.x_train_, **kw)


class LinearNIFKernelOracle(FermionicPQCKernelOracle):
    """linear_nif_kernel.py:107-127: ``h[triu] = act(W x + b)``, ``h[tril] = -h^T``, ``R = matrix_exp(h)``; the Gram is
    the NIFKernel one (per-pair ``R_i R_j^T`` through the lookup-table strategy).  The encoder weights are the caller's
    (the reference draws them from torch's default Linear initialiser)."""

    def __init__(self, n_qubits: int, weight: np.ndarray, bias: Optional[np.ndarray] = None,
                 activation: str = "Identity", gram_batch_size: int = 10_000):
        super().__init__(n_qubits, gram_batch_size=gram_batch_size)
        self.weight = np.asarray(weight, dtype=np.float64)
        self.bias = None if bias is None else np.asarray(bias, dtype=np.float64)
        self.activation = {"Identity": lambda v: v, "Tanh": np.tanh, "ReLU": lambda v: np.maximum(v, 0.0)}[activation]

    def fit(self, x_train, y_train=None):
        self.x_train_ = np.asarray(x_train)
        return self

    def hamiltonian(self, x: np.ndarray) -> np.ndarray:
        x = np.asarray(x, dtype=np.float64).reshape(len(x), -1)
        out = x @ self.weight.T + (0.0 if self.bias is None else self.bias)
        d = 2 * self.n_qubits
        h = np.zeros((len(x), d, d))
        iu = np.triu_indices(d, k=1)
        h[:, iu[0], iu[1]] = self.activation(out)
        il = np.tril_indices(d, k=-1)
        h[:, il[0], il[1]] = -1.0 * h[:, il[1], il[0]]
        return h

    def x_to_sptm(self, x: np.ndarray) -> np.ndarray:
        from scipy.linalg import expm

        return np.stack([expm(hi) for hi in self.hamiltonian(x)])


def indices_batch_generator(shape: Tuple[int, int], batch_size: int = 32) -> Iterator[np.ndarray]:
    """gram_matrix.py:177-201 (vectorised; same cells, same row-major order)."""
    n0, n1 = shape
    ii, jj = np.meshgrid(np.arange(n0), np.arange(n1), indexing="ij")
    mask = (jj > ii) if n0 < n1 else (ii > jj)
    idx = np.stack([ii[mask], jj[mask]], axis=1)
    for s in range(0, len(idx), batch_size):
        yield idx[s:s + batch_size]


class GramMatrixOracle:
    """gram_matrix.py:12-209 without the memmap: float32 storage, computed triangle mirrored, diagonal := 1."""

    def __init__(self, shape, initial_value=0.0, dtype=np.float32):
        self.shape = tuple(shape)
        self.data = np.full(self.shape, initial_value, dtype=dtype)

    def apply_(self, func, batch_size=32, symmetrize=True):
        for indices in indices_batch_generator(self.shape, batch_size):
            self.data[indices[:, 0], indices[:, 1]] = func(indices)
        if symmetrize:
            self.symmetrize_()
        return self

    def symmetrize_(self):
        if self.shape[0] < self.shape[1]:
            idx = np.tril_indices(n=self.shape[0], m=self.shape[1], k=-1)
        else:
            idx = np.triu_indices(n=self.shape[0], m=self.shape[1], k=1)
        self.data[idx[0], idx[1]] = self.data[idx[1], idx[0]]
        np.fill_diagonal(self.data, 1.0)
        return self


class NystroemOracle:
    """nystroem.py:58-123."""

    def __init__(self, kernel: FermionicPQCKernelOracle, n_components=100, random_state=None):
        self.kernel = kernel
        self.n_components = n_components
        self.random_state = random_state

    def fit(self, x, y=None):
        from scipy.linalg import svd
        from sklearn.utils import check_random_state

        self.kernel.fit(x, y)
        rnd = check_random_state(self.random_state)
        n_samples = x.shape[0]
        n_components = min(n_samples, self.n_components)
        inds = rnd.permutation(n_samples)
        basis_inds = inds[:n_components]
        basis = x[basis_inds]
        basis_kernel = self.kernel.gram(basis)
        u, s, v = svd(basis_kernel)
        s = np.maximum(s, 1e-12)
        self.normalization_ = np.dot(u / np.sqrt(s), v)
        self.components_ = basis
        self.component_indices_ = basis_inds
        return self

    def transform(self, x):
        embedded = self.kernel.gram(x, self.components_)
        return np.dot(embedded, self.normalization_.T)
