"""Majorana covariance matrices of product states and their evolution.  TEST INFRASTRUCTURE.

Restates (relative to /root/reference/src/matchcake):
* ``ProductState.from_basis_state`` / ``covariance_matrix`` .. operations/state_preparation/product_state.py:46-79,141-245
* ``displacement_vector`` / ``extended_covariance_matrix`` / ``sptm_lift`` ..
  devices/expval_strategies/m_pfaffian/_extended_covariance.py:12-126
* evolution ``Lambda = R^T Lambda_0 R`` .. devices/nif_device.py:1118-1131 and :1134-1164 (extended)
"""

from __future__ import annotations

import numpy as np


def amplitudes_from_basis_state(bits) -> np.ndarray:
    """product_state.py:46-79: |0> -> [1,0], |1> -> [0,1]; (..., n) -> (..., n, 2) complex."""
    bits = np.asarray(bits).astype(int)
    return np.eye(2, dtype=complex)[bits]


def bloch(amplitudes: np.ndarray):
    """product_state.py:190-193."""
    psi = np.asarray(amplitudes, dtype=complex)
    alpha, beta = psi[..., 0], psi[..., 1]
    x = 2.0 * np.real(np.conj(alpha) * beta)
    y = 2.0 * np.imag(np.conj(alpha) * beta)
    z = np.abs(alpha) ** 2 - np.abs(beta) ** 2
    return x, y, z


def covariance_matrix(amplitudes: np.ndarray) -> np.ndarray:
    """``ProductState.covariance_matrix`` (product_state.py:141-245); amplitudes (..., n, 2) -> (..., 2n, 2n)."""
    x, y, z = bloch(amplitudes)
    n = x.shape[-1]
    batch = x.shape[:-1]
    cov = np.zeros(batch + (2 * n, 2 * n))
    diag = np.arange(n)
    cov[..., 2 * diag, 2 * diag + 1] = -z
    # parity-string factors p_{jk} = prod_{j<l<k} z_l  (:214-225)
    p_mat = np.ones(batch + (n, n))
    for k in range(n):
        for j in range(k - 1):
            p_mat[..., j, k] = np.prod(z[..., j + 1:k], axis=-1)
    rows, cols = np.triu_indices(n, k=1)
    p = p_mat[..., rows, cols]
    cov[..., 2 * rows, 2 * cols] = +y[..., rows] * p * x[..., cols]
    cov[..., 2 * rows, 2 * cols + 1] = +y[..., rows] * p * y[..., cols]
    cov[..., 2 * rows + 1, 2 * cols] = -x[..., rows] * p * x[..., cols]
    cov[..., 2 * rows + 1, 2 * cols + 1] = -x[..., rows] * p * y[..., cols]
    return cov - np.swapaxes(cov, -1, -2)


def displacement_vector(amplitudes: np.ndarray) -> np.ndarray:
    """_extended_covariance.py:12-65: d[2k] = (prod_{l<k} z_l) x_k, d[2k+1] = (prod_{l<k} z_l) y_k."""
    x, y, z = bloch(amplitudes)
    n = x.shape[-1]
    z_prod = np.concatenate([np.ones(x.shape[:-1] + (1,)), np.cumprod(z[..., :-1], axis=-1)], axis=-1) if n > 1 \
        else np.ones_like(x)
    d = np.zeros(x.shape[:-1] + (2 * n,))
    d[..., 0::2] = z_prod * x
    d[..., 1::2] = z_prod * y
    return d


def extended_covariance_matrix(cov: np.ndarray, d: np.ndarray) -> np.ndarray:
    """_extended_covariance.py:68-98: [[cov, d], [-d^T, 0]]."""
    cov = np.asarray(cov, dtype=np.float64)
    d = np.asarray(d, dtype=np.float64)
    n2 = cov.shape[-1]
    out = np.zeros(cov.shape[:-2] + (n2 + 1, n2 + 1))
    out[..., :n2, :n2] = cov
    out[..., :n2, n2] = d
    out[..., n2, :n2] = -d
    return out


def sptm_lift(r: np.ndarray) -> np.ndarray:
    """_extended_covariance.py:101-126: R (+) 1."""
    r = np.asarray(r, dtype=np.float64)
    n2 = r.shape[-1]
    out = np.zeros(r.shape[:-2] + (n2 + 1, n2 + 1))
    out[..., :n2, :n2] = r
    out[..., n2, n2] = 1.0
    return out


def evolve(cov0: np.ndarray, r: np.ndarray) -> np.ndarray:
    """nif_device.py:1131: einsum("...ij,...ik,...kl->...jl", R, cov0, R) = R^T cov0 R."""
    return np.einsum("...ij,...ik,...kl->...jl", r, cov0, r)


def evolved_extended_covariance(amplitudes: np.ndarray, r: np.ndarray) -> np.ndarray:
    """nif_device.py:1134-1164."""
    cov0 = covariance_matrix(amplitudes)
    d0 = displacement_vector(amplitudes)
    return evolve(extended_covariance_matrix(cov0, d0), sptm_lift(r))
