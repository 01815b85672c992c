"""Pauli word -> sorted Majorana index set + phase (host-side integer bookkeeping, bit exact).

Mirrors the interface of the reference's ``matchcake.utils.JordanWigner``
(/root/reference/src/matchcake/utils/jordan_wigner.py:21-140) and ``extend_majorana_indices``
(devices/expval_strategies/m_pfaffian/m_pfaffian_expval_strategy.py:21-38).  Convention (0-based):
``c_{2k} = Z_0..Z_{k-1} X_k``, ``c_{2k+1} = Z_0..Z_{k-1} Y_k``; ``Z_k = -i c_{2k} c_{2k+1}``.

Here the index set is built as a bit mask per Pauli factor and merged with a parity count instead of the
reference's insertion sort over an explicit list; the result (sorted indices, +-1/+-i phase) is identical and is
pinned against the reference's table and brute-force tests (tests/test_host_logic.py).
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple, Union

import numpy as np

_PHASES = (1 + 0j, 0 + 1j, -1 + 0j, 0 - 1j)  # i^q


class JordanWigner:
    def __init__(self, n_qubits: int):
        self.n_qubits = int(n_qubits)

    @staticmethod
    def majorana_x_index(qubit: int) -> int:
        return 2 * qubit

    @staticmethod
    def majorana_y_index(qubit: int) -> int:
        return 2 * qubit + 1

    @staticmethod
    def qubit_of_majorana(mu: int) -> int:
        return mu // 2

    @staticmethod
    def is_x_type(mu: int) -> bool:
        return mu % 2 == 0

    @staticmethod
    def _factor(char: str, k: int) -> Tuple[int, int]:
        """(bit mask of Majorana indices in ascending order, power q of i for the phase i^q)."""
        if char == "I":
            return 0, 0
        if char == "Z":
            return 0b11 << (2 * k), 3  # -i
        if char in ("X", "Y"):
            string = (1 << (2 * k)) - 1  # c_0 .. c_{2k-1}, i.e. Z_0..Z_{k-1}, each contributing (-i)
            own = 1 << (2 * k + (1 if char == "Y" else 0))
            return string | own, (3 * k) % 4
        raise ValueError(f"Unknown Pauli character: {char!r}")

    @staticmethod
    def _merge(acc: int, new: int) -> Tuple[int, int]:
        """Product of two ascending Majorana monomials given as masks: (mask of the product, sign exponent).

        Moving each factor of ``new`` leftwards into place crosses the factors of ``acc`` with a larger index;
        equal indices cancel (c^2 = 1) without a sign because they become adjacent first."""
        swaps = 0
        m = new
        while m:
            low = m & -m
            idx = low.bit_length() - 1
            swaps += bin(acc >> (idx + 1)).count("1")
            acc ^= low
            m ^= low
        return acc, swaps & 1

    def pauli_to_majorana(self, pauli: Union[str, Dict], wires: Optional[Sequence] = None) -> Tuple[np.ndarray, complex]:
        """Return (sorted_indices, phase) such that ``P = phase * c_{mu_1} ... c_{mu_m}``."""
        if isinstance(pauli, dict) or hasattr(pauli, "keys"):
            if wires is None:
                wires = sorted(pauli.keys())
            pairs = [(pauli.get(w, "I"), w) for w in wires]
        else:
            if wires is None:
                raise ValueError("wires must be provided when pauli is a string")
            assert len(pauli) == len(wires), f"len(pauli)={len(pauli)} != len(wires)={len(wires)}"
            pairs = list(zip(pauli, wires))
        wire_to_qubit = {w: k for k, w in enumerate(wires)}
        acc, q = 0, 0
        for char, wire in pairs:
            mask, dq = self._factor(char, wire_to_qubit[wire])
            acc, s = self._merge(acc, mask)
            q = (q + dq + 2 * s) % 4
        idx = [i for i in range(acc.bit_length()) if (acc >> i) & 1]
        return np.array(idx, dtype=int), _PHASES[q]


def extend_majorana_indices(mu: np.ndarray, alpha: complex, parity_index: int) -> Tuple[np.ndarray, complex]:
    """Odd rank -> append the parity index 2n and multiply the phase by i^rank (-1)^{rank(rank-1)/2}."""
    rank = len(mu)
    if rank % 2 == 0:
        return mu, alpha
    q = (rank + 2 * ((rank * (rank - 1) // 2) % 2)) % 4
    return np.concatenate([mu, [parity_index]]).astype(int), alpha * _PHASES[q]
