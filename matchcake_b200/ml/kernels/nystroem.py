"""``Nystroem`` kernel approximation (mirrors /root/reference/src/matchcake/ml/kernels/nystroem.py:11-136).

The two Gram builds (landmark square Gram in ``fit``, rectangular ``K(X, landmarks)`` in ``transform``) run on the
GPU through the kernel; the SVD of the landmark Gram and the final projection are the reference's numpy/scipy
tail (not part of the metric)."""
from __future__ import annotations

import numpy as np
import torch
from scipy.linalg import svd
from sklearn.kernel_approximation import Nystroem as sk_Nystroem
from sklearn.utils import check_random_state
from sklearn.utils.validation import check_is_fitted


def _to_numpy(x):
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


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
        basis_kernel = _to_numpy(self.kernel(basis))
        U, S, V = svd(basis_kernel)
        S = np.maximum(S, 1e-12)
        self.normalization_ = np.dot(U / np.sqrt(S), V)
        self.components_ = basis
        self.component_indices_ = basis_inds
        self._n_features_out = n_components
        return self

    def transform(self, X):
        check_is_fitted(self)
        embedded = _to_numpy(self.kernel(X, self.components_))
        return np.dot(embedded, self.normalization_.T)

    def freeze(self):
        self.kernel.freeze()
        return self
