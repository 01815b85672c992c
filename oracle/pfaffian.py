"""Pfaffian utilities.  TEST INFRASTRUCTURE.

Restates ``utils/_pfaffian.py:17-119`` of the reference, whose arithmetic lives in the third-party
wheel ``torchpfaffian==0.0.3`` (``pyproject.toml:36``, source NOT under /root/reference):

* ``pfaffian(sign=False)`` -> ``torch_pfaffian.PfaffianDet``: ``sqrt(|det A| + EPSILON)`` with the global
  ``EPSILON = 1e-32`` (``_pfaffian.py:81,116-118``).  Whether the floor is added or clamped is not pinned
  by any reference test (they assert 1e-5); we restate the additive form, which reproduces the
  "exact zeros come back as 1e-16" behaviour.
* ``pfaffian(sign=True)`` -> ``torch_pfaffian.pfaffian(sign=True)``: the docstring names the algorithm
  (Parlett-Reid, ``_pfaffian.py:22-24``); restated here from its published form (M. Wimmer,
  "Efficient numerical computation of the Pfaffian for dense and banded skew-symmetric matrices",
  ACM TOMS 38 (2012), Algorithm ``pfaffian_LTL`` with partial pivoting).
* ``sector_pfaffian_features`` -> ``_pfaffian.py:39-75``.
"""

from __future__ import annotations

import numpy as np

EPSILON = 1e-32  # _pfaffian.py:81


def pfaffian_det(matrix: np.ndarray, epsilon: float = EPSILON) -> np.ndarray:
    """``sqrt(|det A| + eps)`` over leading batch axes; accepts real or complex input (_pfaffian.py:110-118)."""
    matrix = np.asarray(matrix)
    n = matrix.shape[-1]
    if n == 0:
        return np.ones(matrix.shape[:-2])
    det = np.linalg.det(matrix)
    return np.sqrt(np.abs(det) + epsilon)


def pfaffian_signed(matrix: np.ndarray) -> np.ndarray:
    """Signed Pfaffian by Parlett-Reid tridiagonalisation with partial pivoting, vectorised over the batch.

    Empty matrix -> 1, odd size -> 0 (pinned by tests/test_utils/test_pfaffian.py:121-127).
    """
    a = np.array(np.real(matrix), dtype=np.float64, copy=True)
    batch_shape = a.shape[:-2]
    n = a.shape[-1]
    if n == 0:
        return np.ones(batch_shape)
    if n % 2 == 1:
        return np.zeros(batch_shape)
    a = a.reshape((-1, n, n))
    nb = a.shape[0]
    rows = np.arange(nb)
    pf = np.ones(nb)
    for k in range(0, n - 1, 2):
        # pivot: largest |A[k, j]|, j > k (row k of the skew matrix), ties -> first
        kp = k + 1 + np.argmax(np.abs(a[:, k, k + 1:]), axis=1)
        swap = kp != k + 1
        if np.any(swap):
            idx = rows[swap]
            p = kp[swap]
            tmp = a[idx, k + 1, :].copy()
            a[idx, k + 1, :] = a[idx, p, :]
            a[idx, p, :] = tmp
            tmp = a[idx, :, k + 1].copy()
            a[idx, :, k + 1] = a[idx, :, p]
            a[idx, :, p] = tmp
            pf[swap] *= -1.0
        piv = a[:, k, k + 1]
        pf *= piv
        if k + 2 < n:
            safe = np.where(piv != 0.0, piv, 1.0)
            tau = a[:, k, k + 2:] / safe[:, None]
            tau[piv == 0.0] = 0.0
            v = a[:, k + 2:, k + 1]
            a[:, k + 2:, k + 2:] += tau[:, :, None] * v[:, None, :] - v[:, :, None] * tau[:, None, :]
    return pf.reshape(batch_shape)


def pfaffian(matrix: np.ndarray, sign: bool = False, epsilon: float = EPSILON) -> np.ndarray:
    """``utils.pfaffian`` (_pfaffian.py:78-119)."""
    matrix = np.asarray(matrix)
    if matrix.shape[-1] % 2 == 1 and not sign:
        # det of an odd skew-symmetric matrix is 0 (up to rounding)
        return pfaffian_det(matrix, epsilon)
    if sign:
        return pfaffian_signed(matrix)
    return pfaffian_det(matrix, epsilon)


def signed_pfaffian(matrix: np.ndarray) -> np.ndarray:
    """_pfaffian.py:17-36."""
    return pfaffian_signed(matrix)


def sector_pfaffian_features(cov_matrix: np.ndarray, index_sets: np.ndarray) -> np.ndarray:
    """_pfaffian.py:39-75: signed Pfaffians of principal sub-blocks, (..., D, D) x (n_terms, m) -> (..., n_terms)."""
    cov = np.asarray(np.real(cov_matrix), dtype=np.float64)
    index_sets = np.asarray(index_sets)
    n_terms, size = index_sets.shape
    sub = cov[..., index_sets[:, :, None], index_sets[:, None, :]]
    if size == 2:
        return sub[..., 0, 1]
    return pfaffian_signed(sub)
