"""The reference's strategy plug-in surface for this path (iii), backed by the B200 kernels.

MatchCake's device picks its arithmetic through three small ABCs whose instances are public attributes of the device:

* ``ProbabilityStrategy``  (devices/probability_strategies/probability_strategy.py:16-148):
  ``__call__(*, state_prep_op, target_binary_states, wires, **kwargs)`` + ``can_execute`` + ``REQUIRES_KWARGS``;
* ``ContractionStrategy``  (devices/contraction_strategies/contraction_strategy.py:14-38):
  ``get_next_operations(op)`` / ``get_reminding()`` / ``reset()``;
* ``ExpvalStrategy``       (devices/expval_strategies/expval_strategy.py:6-22):
  ``__call__(state_prep_op, observable, **kwargs)``.

The classes below keep those names, argument meanings and errors, so that a MatchCake maintainer (INTEGRATION.md
section 3) can register them as they are; each one is a thin adapter onto ``NonInteractingFermionicDevice``'s own
GPU read-outs, i.e. onto the kernels the parity tests exercise.  Nothing here computes on the host.
"""
from __future__ import annotations

from abc import ABC
from typing import Any, List, Optional

from .operations import BasisState, Operation, ProductState, SingleParticleTransitionMatrixOperation, _wires_tuple


# ------------------------------------------------------------------------------------------------ probabilities
class ProbabilityStrategy(ABC):
    NAME: str = "ProbabilityStrategy"
    REQUIRES_KWARGS: List[str] = []

    def check_required_kwargs(self, kwargs: dict) -> "ProbabilityStrategy":
        for kwarg in self.REQUIRES_KWARGS:
            if kwarg not in kwargs:
                raise ValueError(f"Missing required keyword argument: {kwarg}")  # probability_strategy.py:113
        return self

    def can_execute(self, state_prep_op) -> bool:
        return True

    def __call__(self, *, state_prep_op, target_binary_states, wires, **kwargs):
        raise NotImplementedError(f"{type(self).__name__} must implement __call__.")


def _device_for(all_wires, state_prep_op, global_sptm):
    from .device import NonInteractingFermionicDevice

    dev = NonInteractingFermionicDevice(wires=list(_wires_tuple(all_wires)), output_device="cuda")
    if state_prep_op is not None:
        if not dev.apply_state_prep(state_prep_op):
            raise NotImplementedError(f"B200 strategies cannot be used with {type(state_prep_op)}")
    dev.global_sptm = global_sptm
    return dev


class B200ProbabilityStrategy(ProbabilityStrategy):
    """p(y_W) = 2^-k |Pf(Lambda(t)|_W + Lambda_y)| for basis and product states (what LookupTable and ProductState
    strategies of the reference compute, lookup_table_strategy.py:19-71 / product_state_strategy.py:138-213).

    ``target_binary_states`` is ``(k,)`` or ``(Bq, k)``, ``wires`` the measured wires (shared, or one row per outcome);
    required kwargs: ``global_sptm`` (``(2N, 2N)`` or ``(B, 2N, 2N)``) and ``all_wires``.  Returns a scalar / ``(B,)``
    for one outcome and ``(Bq,)`` / ``(Bq, B)`` for a batch, like the device's ``get_states_probability``."""

    NAME = "B200"
    REQUIRES_KWARGS = ["global_sptm", "all_wires"]

    def can_execute(self, state_prep_op) -> bool:
        return isinstance(state_prep_op, (BasisState, ProductState)) or type(state_prep_op).__name__ == "BasisState"

    def __call__(self, *, state_prep_op, target_binary_states, wires, **kwargs):
        self.check_required_kwargs(kwargs)
        dev = _device_for(kwargs["all_wires"], state_prep_op, kwargs["global_sptm"])
        return dev.get_states_probability(target_binary_states, wires)


# ------------------------------------------------------------------------------------------------ expectation values
class ExpvalStrategy(ABC):
    NAME: str = "ExpvalStrategy"

    def can_execute(self, state_prep_op, observable) -> bool:
        return True

    def __call__(self, state_prep_op, observable, **kwargs):
        raise NotImplementedError(f"{type(self).__name__} must implement __call__.")


class B200PfaffianExpvalStrategy(ExpvalStrategy):
    """Pauli words / sums and projectors through the sector-Pfaffian read-out (m_pfaffian_expval_strategy.py:41-306).
    Required kwargs: ``global_sptm`` and ``all_wires`` (the reference passes the same two)."""

    NAME = "B200Pfaffian"
    REQUIRES_KWARGS = ["global_sptm", "all_wires"]

    def __call__(self, state_prep_op, observable, **kwargs):
        for kwarg in self.REQUIRES_KWARGS:
            if kwarg not in kwargs:
                raise ValueError(f"Missing required keyword argument: {kwarg}")
        dev = _device_for(kwargs["all_wires"], state_prep_op, kwargs["global_sptm"])
        return dev.exact_expval(observable)


# ------------------------------------------------------------------------------------------------ contraction
class ContractionStrategy(ABC):
    NAME: str = "ContractionStrategy"
    ALLOWED_GATE_CLASSES: tuple = (Operation,)

    def __init__(self):
        self.p_bar = None
        self.show_progress = False

    def reset(self) -> None:
        pass

    def get_next_operations(self, operation) -> List[Any]:
        raise NotImplementedError

    def get_reminding(self) -> Optional[Any]:
        raise NotImplementedError

    def __call__(self, operations):
        """contraction_strategy.py:40-62: stream the operations through the strategy and collect what it emits."""
        self.reset()
        out: List[Any] = []
        for op in operations:
            out.extend(o for o in self.get_next_operations(op) if o is not None)
        rem = self.get_reminding()
        if rem is not None:
            out.append(rem)
        return out


class B200ContractionStrategy(ContractionStrategy):
    """Collects the whole operation stream and emits ONE ``SingleParticleTransitionMatrixOperation``: the product is
    formed on the GPU by the structured chain kernel (``mcb200_circuit_columns``) instead of one dense padded product
    per group (neighbours_strategy.py:13-26 -> single_particle_transition_matrix.py:228-271,379-402)."""

    NAME = "B200"

    def __init__(self, wires=None):
        super().__init__()
        self.wires = None if wires is None else _wires_tuple(wires)
        self._ops: List[Any] = []

    def reset(self) -> None:
        self._ops = []

    def get_next_operations(self, operation) -> List[Any]:
        self._ops.append(operation)   # nothing is applied eagerly
        return []

    def get_reminding(self) -> Optional[SingleParticleTransitionMatrixOperation]:
        if not self._ops:
            return None
        from .device import NonInteractingFermionicDevice

        wires = self.wires
        if wires is None:
            wires = tuple(sorted(set(w for op in self._ops for w in _wires_tuple(op.wires))))
        if len(wires) < 2:
            raise AssertionError("At least two wires are required for this device.")
        dev = NonInteractingFermionicDevice(wires=list(wires), output_device="cuda")
        dev.apply_generator(self._ops)
        out = dev.global_sptm
        self._ops = []
        return out


def get_probability_strategy(name) -> ProbabilityStrategy:
    """Name -> instance, mirroring devices/probability_strategies/__init__.py (unknown names raise ValueError)."""
    if isinstance(name, ProbabilityStrategy):
        return name
    table = {"b200": B200ProbabilityStrategy, "lookuptable": B200ProbabilityStrategy,
             "productstate": B200ProbabilityStrategy}
    key = str(name).lower().replace("_", "").replace(" ", "")
    if key not in table:
        raise ValueError(f"Unknown probability strategy {name!r}; available: {sorted(table)}")
    return table[key]()


__all__ = ["ProbabilityStrategy", "B200ProbabilityStrategy", "ExpvalStrategy", "B200PfaffianExpvalStrategy",
           "ContractionStrategy", "B200ContractionStrategy", "get_probability_strategy"]
