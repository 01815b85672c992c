"""matchcake_b200 -- B200-native matchgate-circuit contraction path behind MatchCake's device / kernel API.

Only the hot path is here: SPTM chaining, covariance conjugation + Majorana gather, batched
Pfaffians, and the kernel Gram-matrix builder, as hand-written sm_100a CUDA behind the C ABI of
``include/mcb200.h``; plus the thin host-side mirror of the reference interface that calls it.
"""
from . import _lib, observables, operations, ops, torch_ops, utils  # noqa: F401
from .device import DeviceError, NIFDevice, NonInteractingFermionicDevice  # noqa: F401
from . import ml, pennylane_shell, star_state, strategies  # noqa: F401

__version__ = "0.1.0"
