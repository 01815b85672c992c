"""Observables accepted by the device: basis-state projectors and Pauli sums.

PennyLane is absent from the build image, so minimal stand-ins with PennyLane's call shapes are provided
(``Projector(state, wires)``, ``X(0) @ Y(1)``, ``Hamiltonian(coeffs, ops)``).  PennyLane objects are accepted by
duck typing wherever these are (``.parameters[0]`` + ``.wires`` for projectors; ``.terms()`` +
``pennylane.pauli.pauli_word_to_string`` for Pauli sums), so a QNode written against the reference keeps working
when PennyLane is installed.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np


class Projector:
    """|state><state| on ``wires`` (qml.Projector / BasisStateProjector)."""

    def __init__(self, state, wires):
        self.parameters = [np.asarray(state).astype(int)]
        self.wires = tuple(wires.tolist() if hasattr(wires, "tolist") else wires)
        if self.parameters[0].shape[-1] != len(self.wires):
            raise ValueError("Projector: state length must match the number of wires")


BasisStateProjector = Projector


class BatchProjector:
    """A batch of projectors (observables/batch_projector.py:6): states (Bq, k), wires (k,) or (Bq, k)."""

    def __init__(self, states, wires):
        self._states = np.asarray(states).astype(int)
        self._wires = np.asarray(wires)

    def get_states(self):
        return self._states

    def get_batch_wires(self):
        return self._wires


class PauliWord:
    """A tensor product of single-qubit Paulis, e.g. ``X(0) @ Y(2)``."""

    def __init__(self, factors: Dict):
        self.factors = dict(factors)

    @property
    def wires(self) -> Tuple:
        return tuple(self.factors.keys())

    def __matmul__(self, other: "PauliWord") -> "PauliWord":
        f = dict(self.factors)
        for w, c in other.factors.items():
            if w in f and f[w] != "I" and c != "I":
                raise ValueError("Pauli words with overlapping non-identity factors are not supported")
            if c != "I" or w not in f:
                f[w] = c
        return PauliWord(f)

    def __rmul__(self, coeff) -> "PauliSum":
        return PauliSum([coeff], [self])

    __mul__ = __rmul__

    def __add__(self, other):
        return PauliSum([1.0], [self]) + other

    def to_string(self, device_wires: Sequence) -> str:
        for w in self.factors:
            if w not in device_wires:
                raise ValueError(f"wire {w} of the observable is not a device wire")
        return "".join(self.factors.get(w, "I") for w in device_wires)

    def terms(self):
        return [1.0], [self]

    def __repr__(self):
        return " @ ".join(f"{c}({w})" for w, c in self.factors.items())


class PauliSum:
    """``sum_t coeffs[t] * words[t]`` (qml.Hamiltonian / qml.ops.LinearCombination)."""

    def __init__(self, coeffs: Sequence, ops: Sequence[PauliWord]):
        if len(coeffs) != len(ops):
            raise ValueError("coeffs and ops must have the same length")
        self.coeffs = list(coeffs)
        self.ops = list(ops)

    @property
    def wires(self) -> Tuple:
        out: List = []
        for o in self.ops:
            for w in o.wires:
                if w not in out:
                    out.append(w)
        return tuple(out)

    def terms(self):
        return self.coeffs, self.ops

    def __add__(self, other):
        if isinstance(other, PauliWord):
            other = PauliSum([1.0], [other])
        return PauliSum(self.coeffs + other.coeffs, self.ops + other.ops)

    def __rmul__(self, c):
        return PauliSum([c * x for x in self.coeffs], self.ops)

    def __repr__(self):
        return " + ".join(f"{c} * ({o})" for c, o in zip(self.coeffs, self.ops))


Hamiltonian = PauliSum


def X(wire) -> PauliWord:
    return PauliWord({wire: "X"})


def Y(wire) -> PauliWord:
    return PauliWord({wire: "Y"})


def Z(wire) -> PauliWord:
    return PauliWord({wire: "Z"})


def I(wire) -> PauliWord:  # noqa: E743
    return PauliWord({wire: "I"})


def pauli_terms(observable, device_wires: Sequence) -> Tuple[List[complex], List[str]]:
    """(coeffs, Pauli strings over ``device_wires``) of our stand-ins or of a PennyLane operator."""
    if isinstance(observable, (PauliWord, PauliSum)):
        coeffs, ops = observable.terms()
        return [complex(c) for c in coeffs], [o.to_string(device_wires) for o in ops]
    try:  # PennyLane operator
        from pennylane.pauli import pauli_word_to_string

        try:
            coeffs, ops = observable.terms()
        except Exception:
            coeffs, ops = [1.0], [observable]
        wire_map = {w: i for i, w in enumerate(device_wires)}
        out = [pauli_word_to_string(o, wire_map=wire_map) for o in ops]
        return [complex(c.item() if hasattr(c, "item") else c) for c in coeffs], out
    except ImportError as e:
        raise TypeError(f"Unsupported observable {type(observable).__name__}") from e
