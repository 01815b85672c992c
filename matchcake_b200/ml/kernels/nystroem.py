"""``Nystroem`` kernel approximation (mirrors /root/reference/src/matchcake/ml/kernels/nystroem.py:11-136).

The two Gram builds (landmark square Gram in ``fit``, rectangular ``K(X, landmarks)`` in ``transform``) run on the
GPU through the kernel.  The tail (SVD of the landmark Gram, ``embedded @ normalization_.T``;
not part of the metric) stays on the device as well when the kernel hands back CUDA tensors -- plain library calls
(cuSOLVER ``gesvd`` / cuBLAS DGEMM through torch) in float64, so an ``(n, 4096)`` embedding never crosses PCIe twice.
``U diag(S^-1/2) V`` does not depend on the sign / basis choices of the SVD, so the result matches scipy's to rounding."""
from __future__ import annotations

import numpy as np
import torch
from sklearn.kernel_approximation import Nystroem as sk_Nystroem
from sklearn.utils import check_random_state
from sklearn.utils.validation import check_is_fitted


def _to_device(x):
    """float64 tensor; stays where the kernel left it (CUDA for the kernels of this package)."""
    if isinstance(x, torch.Tensor):
        return x.detach().to(torch.float64)
    return torch.as_tensor(np.asarray(x), dtype=torch.float64)


class Nystroem(sk_Nystroem):
    def __init__(self, kernel, *, n_components=100, random_state=None):
        super().__init__(kernel=kernel, n_components=n_components, random_state=random_state)

    def fit(self, X, y=None):
        self.kernel.fit(X, y)
        rnd = check_random_state(self.random_state)
        n_samples = X.shape[0]
        n_components = min(n_samples, self.n_components)
        inds = rnd.permutation(n_samples)
        basis_inds = inds[:n_components]
        basis = X[basis_inds]
        basis_kernel = _to_device(self.kernel(basis))
        U, S, V = torch.linalg.svd(basis_kernel)                      # nystroem.py:96
        S = torch.clamp(S, min=1e-12)                                 # nystroem.py:97
        self._normalization_dev = (U / torch.sqrt(S)) @ V             # nystroem.py:98
        self.normalization_ = self._normalization_dev.cpu().numpy()
        self.components_ = basis
        self.component_indices_ = basis_inds
        self._n_features_out = n_components
        return self

    def _norm_on(self, device) -> torch.Tensor:
        norm = getattr(self, "_normalization_dev", None)
        if norm is None or norm.device != device:                     # e.g. after unpickling / sklearn.clone
            norm = self._normalization_dev = torch.as_tensor(self.normalization_, dtype=torch.float64, device=device)
        return norm

    def transform(self, X):
        check_is_fitted(self)
        embedded = _to_device(self.kernel(X, self.components_))
        return (embedded @ self._norm_on(embedded.device).T).cpu().numpy()   # nystroem.py:123

    def transform_sharded(self, X):
        """Multi-GPU transform that never assembles ``K(X, landmarks)``: every rank builds its row block of the
        rectangular Gram, applies ``@ normalization_.T`` to it on its own device and returns
        ``(embedding_rows (CUDA float64), (row_begin, row_end))`` -- the all-gather of the (n, n_components) matrix is
        left to the caller ("leave sharded and all-gather only if the caller needs it on one device").
        Single process: the whole embedding with range (0, n)."""
        check_is_fitted(self)
        res = self.kernel(X, self.components_, gather=False)
        if isinstance(res, tuple):
            local, rng = res
        else:
            local, rng = res, (0, len(X))
        local = _to_device(local)
        return local @ self._norm_on(local.device).T, rng

    def freeze(self):
        self.kernel.freeze()
        return self
