"""Observable read-out from the global SPTM.  TEST INFRASTRUCTURE.

Restates (relative to /root/reference/src/matchcake):
* ``LookupTableStrategy.__call__`` ............... devices/probability_strategies/lookup_table_strategy.py:19-71
* ``ProductStateProbabilityStrategy`` ............ devices/probability_strategies/product_state_strategy.py:67-213
* dispatcher (LookupTable first for basis states). devices/probability_strategies/probability_func_dispatcher.py:33-66,
                                                   devices/nif_device.py:229-234
* ``states_to_binary`` / ``analytic_probability``  devices/nif_device.py:151-165,430-479
* Jordan-Wigner bookkeeping ...................... utils/jordan_wigner.py:13-140
* ``extend_majorana_indices`` / m-Pfaffian expval  devices/expval_strategies/m_pfaffian/m_pfaffian_expval_strategy.py:21-38,113-158,270-306
"""

from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np

from . import covariance as cov_mod
from . import lookup_table as lt
from .pfaffian import pfaffian_det, sector_pfaffian_features


# --------------------------------------------------------------------------- probabilities
def build_lambda_y(target_binary_state, n_wires: int) -> np.ndarray:
    """product_state_strategy.py:67-92."""
    bits = np.asarray(target_binary_state).astype(int)
    single = bits.ndim == 1
    bits = np.atleast_2d(bits)
    b, k = bits.shape
    lam = np.zeros((b, 2 * k, 2 * k))
    j = np.arange(k)
    vals = -((-1.0) ** bits)
    lam[:, 2 * j, 2 * j + 1] = vals
    lam[:, 2 * j + 1, 2 * j] = -vals
    return lam[0] if single else lam


def majorana_indices_of_wires(wire_indices) -> np.ndarray:
    """product_state_strategy.py:108-110."""
    w = np.asarray(wire_indices)
    return np.stack([2 * w, 2 * w + 1], axis=-1).reshape(w.shape[:-1] + (-1,))


def probs_product_state(cov_t: np.ndarray, target_binary_states, wire_indices) -> np.ndarray:
    """p(y_W) = 2^-k sqrt|det(Lambda|_W + Lambda_y)|  (product_state_strategy.py:138-213).

    ``cov_t`` (..., d, d); ``target_binary_states`` (k,) or (Bq,k); ``wire_indices`` (k,) shared or (Bq,k).
    Returns shape () / (...) for a single outcome, (Bq, ...) for a batch."""
    target = np.asarray(target_binary_states).astype(int)
    single = target.ndim == 1
    k = target.shape[-1]
    w = np.asarray(wire_indices)
    if single or w.ndim == 1 or np.all(w == w[0]):
        wi = w if w.ndim == 1 else w[0]
        mi = majorana_indices_of_wires(wi)
        sub = cov_t[..., mi[:, None], mi[None, :]]
        lam_y = build_lambda_y(target, k)
        if not single:
            lam_y = lam_y.reshape(lam_y.shape[:1] + (1,) * (sub.ndim - 2) + lam_y.shape[1:])
        combined = sub + lam_y
    else:
        mi = majorana_indices_of_wires(w)  # (Bq, 2k)
        sub = cov_t[..., mi[:, :, None], mi[:, None, :]]  # (..., Bq, 2k, 2k)
        combined = sub + build_lambda_y(target, k)
    return (1.0 / 2 ** k) * np.real(pfaffian_det(combined))


def probs_lookup_table(r: np.ndarray, system_state, target_binary_states, wire_indices) -> np.ndarray:
    """Reference default for basis-state inputs: sqrt|det obs| of the complex lookup-table observable
    (lookup_table_strategy.py:19-71).  Returns () / (B,) for a single outcome or (Bq,) / (Bq, B)."""
    target = np.asarray(target_binary_states).astype(int)
    single = target.ndim == 1
    w = np.asarray(wire_indices)
    if not single and w.ndim == 1:
        w = np.broadcast_to(w, target.shape)
    t = lt.transition_matrix(r)
    obs = lt.observable_matrix(t, system_state, target, w)
    if np.asarray(r).ndim == 2:
        obs = obs[:, 0]
    if single:
        obs = obs[0]
    return np.real(pfaffian_det(obs))


def states_to_binary(samples: np.ndarray, num_wires: int) -> np.ndarray:
    """nif_device.py:151-165 (MSB first)."""
    powers = 1 << np.arange(num_wires, dtype=np.int64)
    return ((samples[..., None] & powers) > 0).astype(np.int64)[..., ::-1]


def states_probability(r: np.ndarray, amplitudes: np.ndarray, target_binary_states, wire_indices) -> np.ndarray:
    """``get_states_probability`` (nif_device.py:742-810): LookupTable when the input is a basis state,
    ProductState formula otherwise."""
    amplitudes = np.asarray(amplitudes, dtype=complex)
    is_basis = bool(np.all((np.abs(amplitudes[..., 0]) < 1e-8) | (np.abs(amplitudes[..., 1]) < 1e-8)))
    if is_basis and amplitudes.ndim == 2:
        bits = (np.abs(amplitudes[..., 1]) > np.abs(amplitudes[..., 0])).astype(int)
        return probs_lookup_table(r, bits, target_binary_states, wire_indices)
    cov_t = cov_mod.evolve(cov_mod.covariance_matrix(amplitudes), r)
    return probs_product_state(cov_t, target_binary_states, wire_indices)


def analytic_probability(r: np.ndarray, amplitudes: np.ndarray, wire_indices: Sequence[int]) -> np.ndarray:
    """``qml.probs(wires)``: all 2^k outcomes, transposed to (B, 2^k), divided by the row sum
    (nif_device.py:465-479)."""
    k = len(wire_indices)
    states = states_to_binary(np.arange(2 ** k), k)
    probs = states_probability(r, amplitudes, states, np.asarray(wire_indices))
    probs = np.transpose(probs)
    return probs / np.sum(probs, axis=-1, keepdims=True)


# --------------------------------------------------------------------------- Pauli observables
_PAULI_TO_MAJORANA = {  # jordan_wigner.py:13-18
    "I": lambda k: ([], 1 + 0j),
    "X": lambda k: ([idx for j in range(k) for idx in (2 * j, 2 * j + 1)] + [2 * k], (-1j) ** k),
    "Y": lambda k: ([idx for j in range(k) for idx in (2 * j, 2 * j + 1)] + [2 * k + 1], (-1j) ** k),
    "Z": lambda k: ([2 * k, 2 * k + 1], -1j),
}


def sort_and_cancel(indices: List[int]) -> Tuple[np.ndarray, complex]:
    """jordan_wigner.py:62-85."""
    arr = list(indices)
    sign = 1
    i = 0
    while i < len(arr):
        j = i
        while j > 0 and arr[j - 1] > arr[j]:
            arr[j - 1], arr[j] = arr[j], arr[j - 1]
            sign *= -1
            j -= 1
        if j > 0 and arr[j - 1] == arr[j]:
            arr.pop(j)
            arr.pop(j - 1)
            i = max(0, i - 1)
        else:
            i += 1
    return np.array(arr, dtype=int), complex(sign)


def pauli_to_majorana(pauli: str, wires: Sequence[int]) -> Tuple[np.ndarray, complex]:
    """jordan_wigner.py:109-140; ``pauli`` is a string over IXYZ, one char per entry of ``wires``."""
    assert len(pauli) == len(wires)
    wire_to_qubit = {w: k for k, w in enumerate(wires)}
    indices: List[int] = []
    phase: complex = 1.0 + 0j
    for char, wire in zip(pauli, wires):
        new_indices, local_phase = _PAULI_TO_MAJORANA[char](wire_to_qubit[wire])
        indices.extend(new_indices)
        phase *= local_phase
    if not indices:
        return np.array([], dtype=int), phase
    sorted_indices, perm_sign = sort_and_cancel(indices)
    return sorted_indices, phase * perm_sign


def extend_majorana_indices(mu: np.ndarray, alpha: complex, parity_index: int):
    """m_pfaffian_expval_strategy.py:21-38."""
    rank = len(mu)
    if rank % 2 == 0:
        return mu, alpha
    extra_phase = (1j) ** rank * (-1) ** ((rank * (rank - 1)) // 2)
    return np.concatenate([mu, [parity_index]]), alpha * extra_phase


def pauli_sum_structure(pauli_strs: Sequence[str], coeffs: Sequence[complex], n_qubits: int) -> Dict[int, tuple]:
    """``_build_structure`` + ``_build_payload`` (m_pfaffian_expval_strategy.py:113-158):
    {sector: (index_sets (n_terms, sector) int, scalar_coeffs (n_terms,) float64)}."""
    wires = list(range(n_qubits))
    by_sector: Dict[int, tuple] = {}
    for term_idx, ps in enumerate(pauli_strs):
        mu, alpha = pauli_to_majorana(ps, wires)
        ext, ext_phase = extend_majorana_indices(mu, alpha, 2 * n_qubits)
        idx_list, phase_list, term_list = by_sector.setdefault(len(ext), ([], [], []))
        idx_list.append(ext)
        phase_list.append(ext_phase)
        term_list.append(term_idx)
    payload = {}
    for sector, (idx_list, phase_list, term_list) in by_sector.items():
        index_sets = np.stack(idx_list).astype(int) if sector > 0 else np.zeros((len(idx_list), 0), dtype=int)
        wick = (-1j) ** (sector // 2)
        c = np.asarray([complex(coeffs[t]) for t in term_list], dtype=complex)
        payload[sector] = (index_sets, np.real(c * np.asarray(phase_list, dtype=complex) * wick).astype(np.float64))
    return payload


def expval_pauli_sum(ext_cov: np.ndarray, pauli_strs: Sequence[str], coeffs: Sequence[complex]) -> np.ndarray:
    """``MPfaffianExpvalStrategy.__call__`` (:270-306): sum_sector pf_values @ scalar_coeffs."""
    n_qubits = (ext_cov.shape[-1] - 1) // 2
    payload = pauli_sum_structure(pauli_strs, coeffs, n_qubits)
    total = np.zeros(ext_cov.shape[:-2])
    for sector, (index_sets, scalar_coeffs) in payload.items():
        if sector == 0:
            pf = np.ones(ext_cov.shape[:-2] + (len(scalar_coeffs),))
        elif sector == 2:
            pf = np.real(ext_cov)[..., index_sets[:, 0], index_sets[:, 1]]
        else:
            pf = sector_pfaffian_features(ext_cov, index_sets)
        total = total + pf @ scalar_coeffs
    return total
