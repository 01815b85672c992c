"""PennyLane ``qml.devices.Device`` entry points of the NIF device, PennyLane imported lazily.

Mirrors /root/reference/src/matchcake/devices/nif_device.py: ``execute`` (:562-578), ``_execute_circuit`` (:1057-1075),
``_process_measurement`` (:1077-1094), ``_stopping_condition`` (:1096-1115), ``preprocess_transforms`` (:864-883),
``setup_execution_config`` (:949-967), ``supports_derivatives`` (:969-989), and the shot-based branches of ``expval``
(:679-703) / ``probability`` (:885-909).

PennyLane is not installed in the build image, so everything here works on duck-typed tapes -- any object with
``.operations``, ``.measurements``, ``.wires`` and ``.shots.total_shots`` -- and measurement records are recognised by
class name (``ExpectationMP`` / ``ProbabilityMP`` / ``SampleMP`` anywhere in the MRO).  With PennyLane installed the same
methods receive real ``QuantumScript`` objects, ``preprocess_transforms`` returns PennyLane's own ``CompilePipeline`` and
:func:`pennylane_device_class` builds a ``qml.devices.Device`` subclass; without it ``preprocess_transforms`` returns a
small host pipeline with the same two stages (decompose to supported operations, pass single measurements through).
The arithmetic is the device's: all results come from the CUDA library.
"""
from __future__ import annotations

import dataclasses
from typing import Any, Callable, Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch


def _mro_names(obj) -> Tuple[str, ...]:
    return tuple(c.__name__ for c in type(obj).__mro__)


def _is_tape(obj) -> bool:
    return hasattr(obj, "operations") and hasattr(obj, "measurements")


@dataclasses.dataclass(frozen=True)
class ExecutionConfig:
    """Stand-in for ``pennylane.devices.ExecutionConfig`` when PennyLane is absent (same field names)."""

    grad_on_execution: Optional[bool] = None
    use_device_gradient: Optional[bool] = None
    use_device_jacobian_product: Optional[bool] = None
    gradient_method: Optional[str] = None
    gradient_keyword_arguments: dict = dataclasses.field(default_factory=dict)
    device_options: dict = dataclasses.field(default_factory=dict)
    interface: Optional[str] = None
    derivative_order: int = 1


def _execution_config_cls():
    try:
        from pennylane.devices import ExecutionConfig as QmlExecutionConfig

        return QmlExecutionConfig
    except ImportError:
        return ExecutionConfig


class HostTape:
    """Minimal tape record (``QuantumScript`` call shape) for callers without PennyLane."""

    @dataclasses.dataclass(frozen=True)
    class Shots:
        total_shots: Optional[int] = None

        def __bool__(self):
            return self.total_shots is not None

    def __init__(self, operations: Sequence, measurements: Sequence, shots: Optional[int] = None):
        self.operations = list(operations)
        self.measurements = list(measurements)
        self.shots = HostTape.Shots(shots)

    @property
    def wires(self) -> Tuple:
        seen: List = []
        for o in list(self.operations) + list(self.measurements):
            for w in getattr(o, "wires", ()) or ():
                if w not in seen:
                    seen.append(w)
        return tuple(seen)


class HostPipeline:
    """What ``preprocess_transforms`` returns without PennyLane: callable ``tapes -> (tapes, postprocessing)`` with the
    stages of the reference pipeline -- (1) decompose every operation the device does not support natively until
    ``stopping_condition`` holds (``strict=False``: an operation without a decomposition is left for the device to
    reject), (2) one tape per measurement when the measurements cannot share an execution (the analytic device can
    evaluate any number of measurements from one covariance, so tapes pass through unchanged)."""

    MAX_DEPTH = 16

    def __init__(self, stopping_condition: Callable[[Any], bool], name: str):
        self.stopping_condition = stopping_condition
        self.name = name
        self.transforms = ["decompose", "split_non_commuting"]

    def __len__(self):
        return len(self.transforms)

    def _decompose(self, op, depth=0) -> List:
        if self.stopping_condition(op) or depth >= self.MAX_DEPTH:
            return [op]
        dec = getattr(op, "decomposition", None)
        if dec is None:
            return [op]
        try:
            parts = list(dec())
        except Exception:
            return [op]
        out: List = []
        for p in parts:
            out.extend(self._decompose(p, depth + 1))
        return out

    def __call__(self, tapes):
        single = _is_tape(tapes)
        tapes = [tapes] if single else list(tapes)
        new = []
        for t in tapes:
            ops: List = []
            for op in t.operations:
                ops.extend(self._decompose(op))
            shots = getattr(getattr(t, "shots", None), "total_shots", None)
            new.append(HostTape(ops, t.measurements, shots))
        return tuple(new), (lambda results: results[0] if single else tuple(results))


class PennyLaneDeviceMixin:
    """Mixed into ``NonInteractingFermionicDevice``; relies on its ``reset`` / ``apply_generator`` / ``expval`` /
    ``probability`` / ``generate_samples`` / ``convert_op_to_supported``."""

    #: names the decomposition stops at (nif_device.py:124-135); operation classes of matchcake_b200.operations are
    #: added at import time by device.py
    _supported_ops = {"BasisEmbedding", "StatePrep", "BasisState", "ProductState", "Identity"}

    # ------------------------------------------------------------------ qml.devices.Device API
    def execute(self, circuits, execution_config=None):
        """One result per circuit (a tuple for an iterable of circuits)."""
        if _is_tape(circuits):
            return self._execute_circuit(circuits)
        return tuple(self._execute_circuit(c) for c in circuits)

    def _execute_circuit(self, circuit) -> Any:
        shots = getattr(getattr(circuit, "shots", None), "total_shots", None)
        self._current_shots = shots
        if self._wires is None:
            wires = tuple(circuit.wires)
            assert len(wires) > 1, "At least two wires are required for this device."
            self._wires = wires
        self.reset()
        self.apply_generator(iter(circuit.operations))
        if self._current_shots is not None:
            self._samples = self.generate_samples()
        results = [self._process_measurement(m) for m in circuit.measurements]
        return results[0] if len(results) == 1 else tuple(results)

    def _process_measurement(self, measurement) -> Any:
        names = _mro_names(measurement)
        if "ExpectationMP" in names:
            return self.expval(measurement.obs)
        if "ProbabilityMP" in names:
            w = getattr(measurement, "wires", None)
            return self.probability(wires=w if (w is not None and len(w)) else self.wires)
        if "SampleMP" in names:
            if getattr(measurement, "obs", None) is None:
                return self._samples
            if hasattr(measurement, "process_samples"):
                return measurement.process_samples(self._samples, wire_order=self.wires)
        raise NotImplementedError(f"Measurement {type(measurement).__name__} is not supported by {self.name}.")

    def _stopping_condition(self, op) -> bool:
        if getattr(op, "name", type(op).__name__) in self._supported_ops:
            return True
        from .device import DeviceError

        try:
            self.convert_op_to_supported(op)
            return True
        except DeviceError:
            return False

    def preprocess_transforms(self, execution_config=None):
        try:
            import pennylane as qml
            from pennylane.devices.preprocess import decompose
            from pennylane.transforms.core import CompilePipeline
        except ImportError:
            return HostPipeline(self._stopping_condition, self.name)
        program = CompilePipeline()
        program.add_transform(decompose, stopping_condition=self._stopping_condition, name=self.name, strict=False)
        program.add_transform(_split_non_commuting_transform(qml))
        return program

    def setup_execution_config(self, config=None, circuit=None):
        """Default the gradient method to torch backpropagation (the adjoint kernels sit behind torch autograd)."""
        if config is None:
            config = _execution_config_cls()()
        if config.gradient_method in {"best", None}:
            config = dataclasses.replace(config, gradient_method="backprop")
        return config

    def supports_derivatives(self, execution_config=None, circuit=None) -> bool:
        if execution_config is None:
            return True
        if execution_config.gradient_method in {"backprop", "best"}:
            if circuit is None:
                return True
            return not circuit.shots
        return False

    # ------------------------------------------------------------------ shot-based estimates (process_samples)
    def _samples_view(self, shot_range, bin_size) -> np.ndarray:
        """Recorded samples (shots, [B,] n) restricted to ``shot_range`` and folded into bins of ``bin_size`` shots:
        (bins, bin_size, [B,] n), following pennylane.measurements.*.process_samples."""
        if self._samples is None:
            self._samples = self.generate_samples()
        s = np.asarray(self._samples)
        if shot_range is not None:
            s = s[slice(*shot_range)]
        if bin_size is not None:
            nb = s.shape[0] // bin_size
            return s[: nb * bin_size].reshape((nb, bin_size) + s.shape[1:])
        return s[None]

    def _sampled_probability(self, wires, shot_range=None, bin_size=None):
        """Relative frequencies of the 2^k outcomes of ``wires`` (MSB first) in the recorded samples
        (ProbabilityMP.process_samples)."""
        widx = [self._wire_index(w) for w in wires]
        s = self._samples_view(shot_range, bin_size)[..., widx]             # (bins, shots, [B,] k)
        k = len(widx)
        idx = (s << np.arange(k - 1, -1, -1)).sum(-1)                        # (bins, shots, [B])
        onehot = idx[..., None] == np.arange(1 << k)                         # (bins, shots, [B,] 2^k)
        p = onehot.mean(axis=1)                                              # (bins, [B,] 2^k)
        p = p[0] if bin_size is None else np.moveaxis(p, 0, -1)
        return self._out(torch.as_tensor(p, dtype=torch.float64))

    def _sampled_expval(self, observable, shot_range=None, bin_size=None):
        """Sample mean of a computational-basis-diagonal observable (ExpectationMP.process_samples): basis-state
        projectors and sums of Pauli Z/I words.  The reference feeds the raw computational-basis samples to PennyLane
        without diagonalising rotations, so only such observables have a meaningful shot estimate there as well."""
        from . import observables as obs_mod
        from .device import DeviceError, _is_projector

        if _is_projector(observable):
            widx = [self._wire_index(w) for w in observable.wires]
            s = self._samples_view(shot_range, bin_size)[..., widx]
            tgt = np.asarray(observable.parameters[0]).astype(int).reshape(-1)
            vals = (s == tgt).all(-1).astype(np.float64)                      # (bins, shots, [B])
        else:
            try:
                coeffs, strings = obs_mod.pauli_terms(observable, self.wires)
            except TypeError as err:
                raise DeviceError(f"The observable {observable} cannot be estimated from samples on a {self.name} "
                                  f"device.") from err
            if any(set(ps) - {"I", "Z"} for ps in strings):
                raise DeviceError(f"Shot-based expectation values need an observable that is diagonal in the "
                                  f"computational basis; got {observable}. Use shots=None for the analytic value.")
            s = self._samples_view(shot_range, bin_size)
            vals = np.zeros(s.shape[:-1], dtype=np.float64)
            for c, ps in zip(coeffs, strings):
                mask = np.array([ch == "Z" for ch in ps])
                vals += np.real(c) * (1.0 - 2.0 * (s[..., mask].sum(-1) & 1))
        e = vals.mean(axis=1)                                                  # (bins, [B])
        e = e[0] if bin_size is None else np.moveaxis(e, 0, -1)
        return self._out(torch.as_tensor(e, dtype=torch.float64))


def _split_non_commuting_transform(qml):
    """``_nif_split_non_commuting`` of the reference (nif_device.py:69-78)."""

    @qml.transform
    def _nif_split_non_commuting(tape):
        ms = tape.measurements
        if len(ms) == 1 and ms[0].obs is not None and type(ms[0].obs).__name__ == "BatchHamiltonian":
            return [tape], lambda results: results[0]
        return qml.transforms.split_non_commuting(tape, grouping_strategy="wires")

    return _nif_split_non_commuting


def pennylane_device_class():
    """``qml.devices.Device`` subclass over the B200 device, built on demand (PennyLane must be importable).  Not
    exercised in the build image (no PennyLane); the methods it inherits are the ones tested with tape stubs."""
    import pennylane as qml

    from .device import NonInteractingFermionicDevice

    class NonInteractingFermionicQubitDevice(NonInteractingFermionicDevice, qml.devices.Device):
        name = "nif.qubit"

        def __init__(self, wires=None, shots=None, **kwargs):
            NonInteractingFermionicDevice.__init__(self, wires=wires, shots=shots, **kwargs)
            self._tracker = qml.Tracker()

        @property
        def tracker(self):
            return self._tracker

    return NonInteractingFermionicQubitDevice
