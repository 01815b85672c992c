"""Module-level utilities with the reference's names (matchcake.utils), backed by the CUDA library.

``pfaffian`` / ``signed_pfaffian`` / ``sector_pfaffian_features`` mirror
/root/reference/src/matchcake/utils/_pfaffian.py:17-119: numpy or torch in, same container type out, any number
of leading batch axes.  The arithmetic runs in libmcb200 (warp-per-matrix Parlett-Reid); inputs that live on the
host are copied to the current CUDA device first -- there is no CPU implementation.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from .. import ops as _ops
from .jordan_wigner import JordanWigner, extend_majorana_indices  # noqa: F401


def _to_device(x, dtype=torch.float64) -> torch.Tensor:
    if isinstance(x, torch.Tensor):
        t = x
    else:
        t = torch.as_tensor(np.ascontiguousarray(np.real(x) if np.iscomplexobj(x) else x))
    if t.is_complex():
        t = t.real
    if not t.is_cuda:
        t = t.to("cuda")
    return t.to(dtype).contiguous()


def _like(result: torch.Tensor, ref):
    """utils/math.py convert_and_cast_like: numpy in -> numpy out, torch in -> torch on the input's device/dtype."""
    if isinstance(ref, torch.Tensor):
        dt = ref.dtype if ref.dtype.is_floating_point else (torch.float32 if ref.dtype == torch.complex64 else torch.float64)
        return result.to(device=ref.device, dtype=dt)
    out = result.cpu().numpy()
    ref_dt = np.asarray(ref).dtype
    if ref_dt == np.float32 or ref_dt == np.complex64:
        out = out.astype(np.float32)
    return out


def pfaffian(matrix, sign: bool = False, epsilon: float = 1e-32, dtype: Optional[torch.dtype] = None):
    """Pfaffian of real skew-symmetric matrices ``(..., 2n, 2n)``: signed (``sign=True``) or magnitude.

    Magnitude mode returns ``sqrt(Pf^2 + epsilon)`` = the reference's ``sqrt(|det A| + EPSILON)`` (PfaffianDet,
    _pfaffian.py:110-118): ``|Pf|`` comes from the Parlett-Reid kernel and the additive floor is applied in the
    kernel's epilogue.  ``epsilon=0`` gives ``|Pf|`` itself.  Complex inputs contribute their real part
    (_pfaffian.py:91-95)."""
    t = _to_device(matrix)
    return _like(_ops.pfaffian(t, signed=bool(sign), epsilon=0.0 if sign else float(epsilon)), matrix)


def signed_pfaffian(matrix, dtype: Optional[torch.dtype] = None):
    return pfaffian(matrix, sign=True, dtype=dtype)


def sector_pfaffian_features(cov_matrix, index_sets, dtype: Optional[torch.dtype] = None):
    """``(..., D, D)`` x ``(n_terms, m)`` -> ``(..., n_terms)`` signed Pfaffians of principal sub-blocks."""
    t = _to_device(cov_matrix)
    idx = np.asarray(index_sets)
    n_terms, m = idx.shape
    batch_shape = t.shape[:-2]
    flat = t.reshape((-1,) + t.shape[-2:])
    feats = _ops.sector_features(flat, torch.as_tensor(idx.astype(np.int32)))
    return _like(feats.reshape(batch_shape + (n_terms,)), cov_matrix)


def make_transition_matrix_from_action_matrix(action_matrix):
    """T = 1/2 (R^T[0::2] + i R^T[1::2]) (utils/__init__.py:352-369); layout helper, not on the CUDA path."""
    if isinstance(action_matrix, torch.Tensor):
        rt = action_matrix.transpose(-1, -2)
        return 0.5 * (rt[..., ::2, :] + 1j * rt[..., 1::2, :])
    rt = np.swapaxes(np.asarray(action_matrix), -1, -2)
    return 0.5 * (rt[..., ::2, :] + 1j * rt[..., 1::2, :])


def binary_string_to_vector(s: str) -> np.ndarray:
    return np.array([int(c) for c in s], dtype=int)


def state_to_binary_string(state: int, n: int) -> str:
    return format(int(state), f"0{n}b")
