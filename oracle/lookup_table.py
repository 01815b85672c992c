"""The reference's default projector read-out: the complex "lookup table" observable.  TEST INFRASTRUCTURE.

Restates (relative to /root/reference/src/matchcake):
* ``make_transition_matrix_from_action_matrix`` .. utils/__init__.py:352-369
* ``get_block_diagonal_matrix`` .................. utils/__init__.py:372-394
* ``decompose_binary_state_into_majorana_indexes`` utils/__init__.py:325-349
* the nine lookup-table items .................... base/lookup_table.py:619-710
* index builders ................................. base/lookup_table.py:379-475
* ``compute_observables_of_target_states`` ....... base/lookup_table.py:132-205
"""

from __future__ import annotations

import numpy as np


def transition_matrix(r: np.ndarray) -> np.ndarray:
    """T = 1/2 (R^T[0::2] + i R^T[1::2]); (..., d, d) -> (..., N, d) complex."""
    rt = np.swapaxes(np.asarray(r), -1, -2)
    return 0.5 * (rt[..., ::2, :] + 1j * rt[..., 1::2, :])


def block_diagonal_matrix(n: int) -> np.ndarray:
    b = np.zeros((2 * n, 2 * n), dtype=complex)
    for i in range(n):
        b[2 * i:2 * i + 2, 2 * i:2 * i + 2] = np.array([[1, 1j], [-1j, 1]])
    return b


def lookup_items(t: np.ndarray):
    """3x3 table of items; row/col class 0 = c_d (T), 1 = c_e (T^*), 2 = c_{2p-1} (identity slot)."""
    n = t.shape[-2]
    b = block_diagonal_matrix(n)
    tc = np.conjugate(t)
    tb = np.einsum("...ij,jk->...ik", t, b)            # TB            (:611-613)
    btt = np.einsum("ij,...kj->...ik", b, t)           # B T^T         (:597-599)
    btd = np.einsum("ij,...kj->...ik", b, tc)          # B T^dagger    (:602-608)
    eye = np.zeros(t.shape[:-2] + (t.shape[-1], t.shape[-1]), dtype=complex)
    eye[..., :, :] = np.eye(t.shape[-1])
    return [
        [np.einsum("...pj,...kj->...pk", tb, t), np.einsum("...pj,...kj->...pk", tb, tc), tb],
        [np.einsum("...pi,...ik->...pk", tc, btt), np.einsum("...pi,...ik->...pk", tc, btd),
         np.einsum("...pi,ij->...pj", tc, b)],
        [btt, btd, eye],
    ]


def observable_matrix(t: np.ndarray, system_state, target_binary_states, indexes_of_target_states) -> np.ndarray:
    """``compute_observables_of_target_states``; returns (Bq, B|1, 2(h+k), 2(h+k)) complex antisymmetric."""
    t = np.asarray(t, dtype=complex)
    batched = t.ndim == 3
    tb_ = t if batched else t[None]
    system_state = np.asarray(system_state).astype(int).flatten()
    target = np.asarray(target_binary_states).astype(int)
    target = target.reshape(-1, target.shape[-1])
    idx_t = np.asarray(indexes_of_target_states).astype(int)
    idx_t = idx_t.reshape(-1, idx_t.shape[-1])
    bq = target.shape[0]

    ket = 2 * np.nonzero(system_state)[0]                       # :325-349
    bra = ket[::-1]
    ket_idx = np.repeat(ket[None, :], bq, axis=0)
    bra_idx = np.repeat(bra[None, :], bq, axis=0)
    unmeasured = np.full((bq, ket.size), 2, dtype=int)
    measured_cls = np.stack([target, 1 - target], axis=-1).reshape(bq, -1)    # :413-439
    lt_idx = np.concatenate([unmeasured, measured_cls, unmeasured], axis=-1)
    measured_pos = np.stack([idx_t, idx_t], axis=-1).reshape(bq, -1)          # :441-475
    maj_idx = np.concatenate([bra_idx, measured_pos, ket_idx], axis=-1)

    size = maj_idx.shape[-1]
    iu = np.triu_indices(size, k=1)
    items = lookup_items(tb_)
    # pad every item to a common (B, dmax, dmax) and stack to (9, B, dmax, dmax)   (:497-536)
    dmax = tb_.shape[-1]
    stacked = np.zeros((3, 3, tb_.shape[0], dmax, dmax), dtype=complex)
    for a in range(3):
        for b in range(3):
            it = items[a][b]
            stacked[a, b, :, :it.shape[-2], :it.shape[-1]] = it
    cls_r, cls_c = lt_idx[:, iu[0]], lt_idx[:, iu[1]]           # (Bq, P)
    pos_r, pos_c = maj_idx[:, iu[0]], maj_idx[:, iu[1]]
    vals = stacked[cls_r, cls_c, :, pos_r, pos_c]               # (Bq, P, B)
    obs = np.zeros((bq, tb_.shape[0], size, size), dtype=complex)
    obs[..., iu[0], iu[1]] = np.transpose(vals, (0, 2, 1))
    obs = obs - np.swapaxes(obs, -1, -2)                        # :198
    return obs
