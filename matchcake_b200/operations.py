"""Gate and state-preparation records with the reference's names and argument meaning.

Mirrors /root/reference/src/matchcake/operations: ``CompRxRx/CompRyRy/CompRzRz`` (comp_rotations.py:237-285),
``fSWAP = FermionicSWAP = CompZX`` (fermionic_swap.py:15-61), generic ``MatchgateOperation``
(matchgate_operation.py:26), ``SingleParticleTransitionMatrixOperation`` and its closed-form subclasses
(single_particle_transition_matrices/), ``SptmAngleEmbedding`` (sptm_angle_embedding.py:17-79) and
``ProductState`` (state_preparation/product_state.py).

These objects only DESCRIBE a circuit: they carry wires and parameter tensors.  The device compiles a stream of
them into one circuit descriptor and the arithmetic (trig, chaining, conjugation, Pfaffians) runs in the CUDA
library.  PennyLane is not required; when it is installed its ``Wires`` are accepted wherever wires are.
"""
from __future__ import annotations

from typing import Any, Iterable, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import ops as _ops

TensorLike = Union[np.ndarray, torch.Tensor, Sequence[float]]


def _wires_tuple(wires) -> Tuple[Any, ...]:
    t = type(wires)
    if t is tuple:
        return wires
    if t is list:
        return tuple(wires)
    if wires is None:
        return ()
    if isinstance(wires, (int, np.integer, str)):
        return (wires,)
    if hasattr(wires, "tolist"):
        return tuple(wires.tolist())
    return tuple(wires)


def make_wires_continuous(wires: Sequence[int]) -> Tuple[int, ...]:
    return tuple(range(min(wires), max(wires) + 1))


def _shape(x) -> Tuple[int, ...]:
    return tuple(x.shape) if hasattr(x, "shape") else tuple(np.shape(x))


def _as_f64_tensor(x, device=None, keep_grad: bool = False) -> torch.Tensor:
    if isinstance(x, torch.Tensor):
        t = x if (keep_grad and x.requires_grad and torch.is_grad_enabled()) else x.detach()
        if t.is_complex():
            t = t.real
    else:
        a = np.asarray(x)
        t = torch.as_tensor(np.ascontiguousarray(a.real if np.iscomplexobj(a) else a))
    t = t.to(torch.float64)
    if device is not None:
        # pinned host memory: asynchronous copy on the current stream (the kernels that consume it run on that stream)
        t = t.to(device, non_blocking=t.device.type == "cpu" and t.is_pinned())
    return t


class Operation:
    """Common base: ``wires`` (tuple of labels), ``name``, optional leading batch axis of the data."""

    num_wires: Optional[int] = None

    def __init__(self, wires=None):
        self.wires = _wires_tuple(wires)

    @property
    def name(self) -> str:
        return type(self).__name__

    @property
    def cs_wires(self) -> Tuple[int, ...]:
        return make_wires_continuous(self.wires)

    @property
    def sorted_wires(self) -> Tuple[int, ...]:
        return tuple(sorted(self.wires))

    @property
    def batch_size(self) -> Optional[int]:
        return None

    def __repr__(self):
        return f"{self.name}(wires={list(self.wires)})"


# ------------------------------------------------------------------------------------------ matchgates
class MatchgateOperation(Operation):
    r"""A two-qubit matchgate ``[[a,0,0,b],[0,w,x,0],[0,y,z,0],[c,0,0,d]]`` on consecutive wires.

    ``matrix``: ``(4,4)`` or ``(B,4,4)`` complex.  The structure and unitarity constraints of
    matchgate_operation.py:382-432 are checked on construction (``check=False`` skips)."""

    num_wires = 2

    def __init__(self, matrix: TensorLike, wires=None, check: bool = True, **kwargs):
        super().__init__(wires)
        if len(self.wires) != 2:
            raise AssertionError(f"MatchgateOperation requires exactly 2 wires, got {len(self.wires)}.")
        if self.wires[1] - self.wires[0] != 1:
            raise AssertionError(f"MatchgateOperation requires consecutive wires, got {self.wires}.")
        m = matrix.detach().cpu().numpy() if isinstance(matrix, torch.Tensor) else np.asarray(matrix)
        self._matrix = m.astype(complex)
        #: the caller's tensor when it takes part in autograd: the SPTM block is then derived from IT (mcb200_matchgate_sptm
        #: and its adjoint kernel), so gradients reach the parameters the matchgate was built from
        self._matrix_t = matrix if (isinstance(matrix, torch.Tensor) and matrix.requires_grad) else None
        if self._matrix.shape[-2:] != (4, 4):
            raise ValueError(f"Matchgate matrix must be (..., 4, 4), got {self._matrix.shape}")
        self.kwargs = kwargs
        if check:
            self._check_is_matchgate()

    def _check_is_matchgate(self, atol: float = 1e-6):
        m = self._matrix
        zero_mask = np.ones((4, 4), dtype=bool)
        for i, j in [(0, 0), (0, 3), (3, 0), (3, 3), (1, 1), (1, 2), (2, 1), (2, 2)]:
            zero_mask[i, j] = False
        if not np.allclose(m[..., zero_mask], 0.0, atol=atol):
            raise ValueError("The matrix does not have the matchgate structure.")
        eye = np.eye(4)
        if not np.allclose(m @ np.conjugate(np.swapaxes(m, -1, -2)), eye, atol=atol):
            raise ValueError("The matchgate is not unitary.")
        det_o = m[..., 0, 0] * m[..., 3, 3] - m[..., 0, 3] * m[..., 3, 0]
        det_i = m[..., 1, 1] * m[..., 2, 2] - m[..., 1, 2] * m[..., 2, 1]
        if not np.allclose(det_o, det_i, atol=atol):
            raise ValueError("The determinants of the outer and inner blocks differ: not a matchgate.")

    def matrix(self) -> np.ndarray:
        return self._matrix

    @property
    def batch_size(self) -> Optional[int]:
        return None if self._matrix.ndim == 2 else self._matrix.shape[0]

    @property
    def single_particle_transition_matrix(self):
        r"""``R_{\mu\nu} = 1/4 Tr(U c_\mu U^\dagger c_\nu)`` (utils/__init__.py:561-596), real (…, 4, 4).  Computed by the
        CUDA library (``mcb200_matchgate_sptm``); raises ``ValueError`` if the imaginary part exceeds 1e-6."""
        if self._matrix_t is not None and torch.is_grad_enabled():
            return _ops.matchgate_sptm(self._matrix_t.to("cuda"))
        return _ops.matchgate_sptm(torch.as_tensor(self._matrix).to("cuda"))

    def to_sptm_operation(self) -> "SingleParticleTransitionMatrixOperation":
        return SingleParticleTransitionMatrixOperation(self.single_particle_transition_matrix, wires=self.wires)

    def __matmul__(self, other):
        if isinstance(other, SingleParticleTransitionMatrixOperation):
            return self.to_sptm_operation() @ other
        if not isinstance(other, MatchgateOperation):
            raise ValueError(f"Cannot multiply MatchgateOperation with {type(other)}")
        if self.wires != other.wires:
            return self.to_sptm_operation() @ other.to_sptm_operation()
        if self._matrix_t is not None or other._matrix_t is not None:   # keep the product inside autograd
            a = self._matrix_t if self._matrix_t is not None else torch.as_tensor(self._matrix)
            b = other._matrix_t if other._matrix_t is not None else torch.as_tensor(other._matrix)
            return MatchgateOperation(a.to(torch.complex128) @ b.to(device=a.device, dtype=torch.complex128),
                                      wires=self.wires, check=False)
        return MatchgateOperation(self._matrix @ other._matrix, wires=self.wires, check=False)


_MAJ2: Optional[np.ndarray] = None


def _majorana2() -> np.ndarray:
    global _MAJ2
    if _MAJ2 is None:
        x = np.array([[0, 1], [1, 0]], dtype=complex)
        y = np.array([[0, -1j], [1j, 0]])
        z = np.diag([1.0 + 0j, -1.0])
        i2 = np.eye(2)
        _MAJ2 = np.stack([np.kron(x, i2), np.kron(y, i2), np.kron(z, x), np.kron(z, y)])
    return _MAJ2


def matchgate_to_sptm_block(u: np.ndarray) -> np.ndarray:
    """(..., 4, 4) complex matchgate -> (..., 4, 4) real SPTM block on the HOST: only for constants that are baked into a
    circuit descriptor once (the Hadamard entangler block of FermionicPQCKernel); batched / runtime matchgates go through
    the ``mcb200_matchgate_sptm`` kernel (``ops.matchgate_sptm``)."""
    c = _majorana2()
    u = np.asarray(u, dtype=complex)
    uc = np.einsum("...ij,mjq->...miq", u, c)
    ucu = np.einsum("...miq,...kq->...mik", uc, np.conjugate(u))
    r = np.einsum("...kij,mji->...km", ucu, c) / 4.0
    if np.max(np.abs(r.imag)) > SingleParticleTransitionMatrixOperation.REAL_PART_ATOL:
        raise ValueError("The single-particle transition matrix must be real (is the gate a matchgate?)")
    return np.ascontiguousarray(r.real)


class CompRotation(MatchgateOperation):
    """``U = M(R_P(theta), R_P(phi))``, ``params = (..., 2) = (theta, phi)`` (comp_rotations.py:83-235)."""

    DIRECTION = "X"
    GATE_KIND = _ops.GATE_RXRX

    def __init__(self, params: TensorLike, wires=None, **kwargs):
        w = self.wires = _wires_tuple(wires)
        if len(w) != 2 or w[1] - w[0] != 1:
            raise AssertionError(f"{self.name} requires 2 consecutive wires, got {self.wires}.")
        shp = params.shape if hasattr(params, "shape") else np.shape(params)
        if len(shp) not in (1, 2) or shp[-1] != 2:
            raise ValueError(f"Invalid shape for the parameters: {tuple(shp)}. Expected (2,) or (batch, 2).")
        self._given_params = params
        self.kwargs = kwargs

    @classmethod
    def random_params(cls, batch_size=None, **kwargs):
        shape = ([batch_size] if batch_size is not None else []) + [2]
        return np.random.default_rng(kwargs.pop("seed", None)).uniform(0, 2 * np.pi, shape)

    @classmethod
    def random(cls, wires, *, batch_size=None, **kwargs):
        return cls(cls.random_params(batch_size=batch_size, **kwargs), wires=wires)

    @property
    def params(self):
        return self._given_params

    @property
    def batch_size(self) -> Optional[int]:
        shp = _shape(self._given_params)
        return shp[0] if len(shp) == 2 else None

    def matrix(self) -> np.ndarray:
        """4x4 complex matchgate (comp_rotations.py:18-80); for inspection, never used by the device."""
        p = _as_f64_tensor(self._given_params).cpu().numpy()
        th, ph = p[..., 0], p[..., 1]

        def rot(a):
            m = np.zeros(a.shape + (2, 2), dtype=complex)
            if self.DIRECTION == "X":
                m[..., 0, 0] = m[..., 1, 1] = np.cos(a / 2)
                m[..., 0, 1] = m[..., 1, 0] = -1j * np.sin(a / 2)
            elif self.DIRECTION == "Y":
                m[..., 0, 0] = m[..., 1, 1] = np.cos(a / 2)
                m[..., 0, 1] = -np.sin(a / 2)
                m[..., 1, 0] = np.sin(a / 2)
            else:
                m[..., 0, 0] = np.exp(-1j * a / 2)
                m[..., 1, 1] = np.exp(1j * a / 2)
            return m

        o, i = rot(th), rot(ph)
        m = np.zeros(p.shape[:-1] + (4, 4), dtype=complex)
        m[..., 0, 0], m[..., 0, 3], m[..., 3, 0], m[..., 3, 3] = o[..., 0, 0], o[..., 0, 1], o[..., 1, 0], o[..., 1, 1]
        m[..., 1, 1], m[..., 1, 2], m[..., 2, 1], m[..., 2, 2] = i[..., 0, 0], i[..., 0, 1], i[..., 1, 0], i[..., 1, 1]
        return m

    @property
    def _matrix(self):  # generic-matchgate view for MatchgateOperation.__matmul__
        return self.matrix()

    def to_sptm_operation(self):
        return _SPTM_OF_ROT[type(self).DIRECTION](self._given_params, wires=self.wires)

    def __repr__(self):
        return f"{self.name}(params shape={_shape(self._given_params)}, wires={list(self.wires)})"


class CompRxRx(CompRotation):
    DIRECTION, GATE_KIND = "X", _ops.GATE_RXRX


class CompRyRy(CompRotation):
    DIRECTION, GATE_KIND = "Y", _ops.GATE_RYRY


class CompRzRz(CompRotation):
    DIRECTION, GATE_KIND = "Z", _ops.GATE_RZRZ


class CompZX(MatchgateOperation):
    """fSWAP = M(Z, X) (fermionic_swap.py:15-61)."""

    def __init__(self, wires=None, **kwargs):
        w = self.wires = _wires_tuple(wires)
        if len(w) != 2 or abs(w[1] - w[0]) != 1:
            raise AssertionError(f"fSWAP as a matchgate requires 2 consecutive wires, got {self.wires}.")
        self.kwargs = kwargs

    @classmethod
    def random(cls, wires, batch_size=None, **kwargs):
        return cls(wires=wires)

    def matrix(self) -> np.ndarray:
        return np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, -1]], dtype=complex)

    @property
    def _matrix(self):
        return self.matrix()

    def to_sptm_operation(self):
        return SptmCompZX(wires=self.wires)


FermionicSWAP = CompZX
fSWAP = CompZX


# ------------------------------------------------------------------------------------------ SPTM ops
class SingleParticleTransitionMatrixOperation(Operation):
    """A real orthogonal ``(2w, 2w)`` / ``(B, 2w, 2w)`` block on a continuous span of ``w`` wires
    (single_particle_transition_matrix.py:21-541)."""

    REAL_PART_ATOL = 1e-6
    GATE_KIND: Optional[int] = None

    def __init__(self, matrix: TensorLike, wires=None, check_real: bool = True, **kwargs):
        super().__init__(wires)
        self._hyperparameters = kwargs
        if isinstance(matrix, torch.Tensor):
            if matrix.is_complex():
                if check_real and matrix.numel() and float(matrix.imag.abs().max()) > self.REAL_PART_ATOL:
                    raise ValueError(
                        f"The single-particle transition matrix must be real, but its maximum absolute imaginary "
                        f"part is {float(matrix.imag.abs().max())}, which exceeds {self.REAL_PART_ATOL}.")
                matrix = matrix.real
            self._mat = matrix
        else:
            m = np.asarray(matrix)
            if np.iscomplexobj(m):
                if check_real and m.size and np.max(np.abs(m.imag)) > self.REAL_PART_ATOL:
                    raise ValueError(
                        f"The single-particle transition matrix must be real, but its maximum absolute imaginary "
                        f"part is {np.max(np.abs(m.imag))}, which exceeds {self.REAL_PART_ATOL}.")
                m = m.real
            self._mat = np.ascontiguousarray(m, dtype=np.float64)
        shp = _shape(self._mat)
        if len(shp) not in (2, 3) or shp[-1] != shp[-2]:
            raise ValueError(f"SPTM must be (2w,2w) or (B,2w,2w), got {shp}")
        if self.wires and shp[-1] != 2 * len(self.cs_wires):
            raise ValueError(f"SPTM of size {shp[-1]} does not match the span of wires {self.wires}")

    # -- constructors ----------------------------------------------------------------------------
    @classmethod
    def from_operation(cls, op, **kwargs) -> "SingleParticleTransitionMatrixOperation":
        if isinstance(op, SingleParticleTransitionMatrixOperation):
            return op
        if hasattr(op, "to_sptm_operation") and callable(op.to_sptm_operation):
            return op.to_sptm_operation()
        if hasattr(op, "single_particle_transition_matrix"):
            return SingleParticleTransitionMatrixOperation(op.single_particle_transition_matrix, wires=op.wires)
        try:
            return MatchgateOperation(op.matrix(), wires=op.wires).to_sptm_operation()
        except Exception as e:
            raise ValueError(f"Cannot convert {type(op)} to {cls.__name__}") from e

    @classmethod
    def from_operations(cls, ops: Iterable, **kwargs) -> Optional["SingleParticleTransitionMatrixOperation"]:
        """Block-diagonal assembly of operations acting on disjoint wires (:228-271)."""
        ops = [cls.from_operation(op) for op in ops]
        if not ops:
            return None
        if len(ops) == 1:
            return ops[0]
        all_wires = make_wires_continuous(sorted(set(w for op in ops for w in op.cs_wires)))
        d = 2 * len(all_wires)
        mats = [_as_f64_tensor(op.matrix()) for op in ops]
        device = next((m.device for m in mats if m.is_cuda), mats[0].device)
        bs = [m.shape[0] for m in mats if m.ndim == 3]
        out = torch.eye(d, dtype=torch.float64, device=device)
        if bs:
            out = out.repeat(bs[0], 1, 1)
        for op, m in zip(ops, mats):
            w0 = all_wires.index(op.sorted_wires[0])
            out[..., 2 * w0:2 * w0 + m.shape[-2], 2 * w0:2 * w0 + m.shape[-1]] = m.to(device)
        return SingleParticleTransitionMatrixOperation(out, wires=all_wires)

    @classmethod
    def random_params(cls, batch_size=None, **kwargs):
        """default_rng(seed).normal upper triangle h -> expm(4 h) (:274-287)."""
        from scipy.linalg import expm

        wires = kwargs.get("wires")
        assert wires is not None, "wires kwarg must be set."
        rng = np.random.default_rng(kwargs.pop("seed", None))
        d = 2 * len(_wires_tuple(wires))
        iu = np.triu_indices(d, k=1)
        shape = [batch_size] if batch_size is not None else []
        rn = rng.normal(size=shape + [iu[0].size])
        m = np.zeros(shape + [d, d], dtype=complex)
        m[..., iu[0], iu[1]] = rn
        m[..., iu[1], iu[0]] = -rn
        return expm(4 * m)

    @classmethod
    def random(cls, wires, batch_size=None, **kwargs):
        return cls(cls.random_params(batch_size=batch_size, wires=wires, **kwargs), wires=wires)

    # -- accessors -------------------------------------------------------------------------------
    def matrix(self, wire_order=None):
        return self._mat

    @property
    def batch_size(self) -> Optional[int]:
        shp = _shape(self._mat)
        return shp[0] if len(shp) == 3 else None

    def pad(self, wires) -> "SingleParticleTransitionMatrixOperation":
        """Embed at offset ``2 * index(wire0)`` of an identity on the continuous span of ``wires`` (:379-402)."""
        cs = make_wires_continuous(_wires_tuple(wires))
        if self.cs_wires == cs:
            return self
        m = _as_f64_tensor(self._mat)
        d = 2 * len(cs)
        out = torch.eye(d, dtype=torch.float64, device=m.device)
        if m.ndim == 3:
            out = out.repeat(m.shape[0], 1, 1)
        w0 = cs.index(self.cs_wires[0])
        out[..., 2 * w0:2 * w0 + m.shape[-2], 2 * w0:2 * w0 + m.shape[-1]] = m
        return SingleParticleTransitionMatrixOperation(out, wires=cs, **self._hyperparameters)

    def __matmul__(self, other):
        """``first @ second`` on the union span, on the GPU (:404-418)."""
        if not isinstance(other, SingleParticleTransitionMatrixOperation):
            if hasattr(other, "to_sptm_operation"):
                other = other.to_sptm_operation()
            else:
                raise ValueError(f"Cannot multiply {self.__class__.__name__} with {type(other)}")
        wires = make_wires_continuous(sorted(set(self.wires) | set(other.wires)))
        a = _as_f64_tensor(self.pad(wires).matrix(), device="cuda")
        b = _as_f64_tensor(other.pad(wires).matrix(), device="cuda")
        return SingleParticleTransitionMatrixOperation(_ops.sptm_matmul(a, b), wires=wires, **self._hyperparameters)

    def adjoint(self) -> "SingleParticleTransitionMatrixOperation":
        m = self._mat
        mt = m.transpose(-1, -2) if isinstance(m, torch.Tensor) else np.swapaxes(m, -1, -2)
        return SingleParticleTransitionMatrixOperation(mt, wires=self.wires, **self._hyperparameters)

    def to_sptm_operation(self):
        return self

    def __repr__(self):
        return f"{self.name}(shape={_shape(self._mat)}, wires={list(self.wires)})"


class _SptmCompRotation(SingleParticleTransitionMatrixOperation):
    GATE_KIND = _ops.GATE_RXRX
    _MATCHGATE = CompRxRx

    def __init__(self, params: TensorLike, wires=None, **kwargs):
        Operation.__init__(self, wires)
        shp = _shape(params)
        if shp[-1] != 2:
            raise ValueError(f"Invalid number of parameters: {shp[-1]}. Expected 2.")
        if len(shp) not in (1, 2):
            raise ValueError(f"Invalid shape for the parameters: {shp}")
        if len(self.wires) != 2 or self.wires[1] - self.wires[0] != 1:
            raise ValueError(f"{self.name} acts on two consecutive wires, got {self.wires}")
        self._given_params = params
        self._hyperparameters = kwargs

    @classmethod
    def random_params(cls, batch_size=None, **kwargs):
        shape = ([batch_size] if batch_size is not None else []) + [2]
        return np.random.default_rng(kwargs.pop("seed", None)).uniform(0, 2 * np.pi, shape)

    @classmethod
    def random(cls, wires, batch_size=None, **kwargs):
        return cls(cls.random_params(batch_size=batch_size, **kwargs), wires=wires)

    @property
    def params(self):
        return self._given_params

    @property
    def batch_size(self) -> Optional[int]:
        shp = _shape(self._given_params)
        return shp[0] if len(shp) == 2 else None

    @property
    def _mat(self):
        """Closed form evaluated by the CUDA circuit kernel on a two-wire circuit."""
        p = _as_f64_tensor(self._given_params, device="cuda")
        single = p.ndim == 1
        circ = _ops.CompiledCircuit(2, [_ops.GateSpec(self.GATE_KIND, 0, 1, 0)], 2)
        r = _ops.circuit_columns(circ, p.reshape(-1, 2))
        return r[0] if single else r

    def to_matchgate(self):
        return self._MATCHGATE(self._given_params, wires=self.wires)


class SptmCompRxRx(_SptmCompRotation):
    GATE_KIND, _MATCHGATE = _ops.GATE_RXRX, CompRxRx


class SptmCompRyRy(_SptmCompRotation):
    GATE_KIND, _MATCHGATE = _ops.GATE_RYRY, CompRyRy


class SptmCompRzRz(_SptmCompRotation):
    GATE_KIND, _MATCHGATE = _ops.GATE_RZRZ, CompRzRz


_SPTM_OF_ROT = {"X": SptmCompRxRx, "Y": SptmCompRyRy, "Z": SptmCompRzRz}


class SptmCompZX(SingleParticleTransitionMatrixOperation):
    """fSWAP as an SPTM; non-adjacent wires allowed; self-adjoint (sptm_fswap.py:7-36)."""

    GATE_KIND = _ops.GATE_FSWAP

    def __init__(self, wires=None, **kwargs):
        Operation.__init__(self, wires)
        if len(self.wires) != 2 or self.wires[0] == self.wires[1]:
            raise ValueError(f"fSWAP acts on two distinct wires, got {self.wires}")
        self._hyperparameters = kwargs

    @classmethod
    def random(cls, wires, batch_size=None, **kwargs):
        return cls(wires=wires)

    @property
    def _mat(self):
        w0, w1 = sorted(self.wires)
        size = 2 * (w1 - w0 + 1)
        m = np.zeros((size, size))
        m[0, size - 2] = m[1, size - 1] = 1.0
        m[np.arange(2, size), np.arange(size - 2)] = 1.0
        return m.T.copy() if w0 != self.wires[0] else m

    def adjoint(self):
        return self

    def to_matchgate(self):
        return CompZX(wires=self.wires)


SptmFSwap = SptmCompZX

ROT = {"X": SptmCompRxRx, "Y": SptmCompRyRy, "Z": SptmCompRzRz}


class SptmAngleEmbedding(Operation):
    """Feature pairs -> rotation blocks on wire pairs (sptm_angle_embedding.py:12-79)."""

    def __init__(self, features: TensorLike, wires, rotations: str = "X", **kwargs):
        features = self.pad_params(features)
        n_features = _shape(features)[-1]
        wires = _wires_tuple(wires)
        if n_features > len(wires):
            raise ValueError(f"Features must be of length {len(wires)} or less; got length {n_features}.")
        self._rotations = rotations.split(",")
        self.features = features
        super().__init__(wires[:n_features])

    @staticmethod
    def pad_params(params):
        n = _shape(params)[-1]
        if n % 2 != 0:
            if isinstance(params, torch.Tensor):
                params = torch.cat([params, torch.zeros_like(params[..., :1])], dim=-1)
            else:
                params = np.concatenate([np.asarray(params), np.zeros_like(np.asarray(params)[..., :1])], axis=-1)
        return params

    def decomposition(self) -> List[SingleParticleTransitionMatrixOperation]:
        f = self.features
        out = []
        for i in range(_shape(f)[-1] // 2):
            p = f[..., 2 * i:2 * i + 2]
            for r in self._rotations:
                out.append(ROT[r](p, wires=[self.wires[2 * i], self.wires[2 * i + 1]]))
        return out


# ------------------------------------------------------------------------------------------ state prep
class BasisState(Operation):
    """Computational-basis state preparation (stands in for ``qml.BasisState``)."""

    def __init__(self, state, wires):
        super().__init__(wires)
        self.parameters = [np.asarray(state).astype(int)]
        if self.parameters[0].shape[-1] != len(self.wires):
            raise ValueError("BasisState: state length must match the number of wires")


class ProductState(Operation):
    """Per-qubit amplitudes ``(n,2)`` (or flat ``(2n,)``, batched ``(B,n,2)``/``(B,2n)``) -- product_state.py:81-139."""

    ATOL = 1e-8

    def __init__(self, state: TensorLike, wires, validate_norm: bool = True):
        super().__init__(wires)
        s = state.detach().cpu().numpy() if isinstance(state, torch.Tensor) else np.asarray(state)
        n = len(self.wires)
        shape = s.shape
        if shape == (2 * n,):
            s = s.reshape(n, 2)
        elif shape == (n, 2):
            pass
        elif len(shape) == 2 and shape[1] == 2 * n:
            s = s.reshape(shape[0], n, 2)
        elif len(shape) == 3 and shape[1:] == (n, 2):
            pass
        else:
            raise ValueError(
                f"ProductState expects a flat state vector of length 2 * n_wires = {2 * n}, an (n_wires, 2) = "
                f"({n}, 2) matrix, or their batched forms (B, {2 * n}) / (B, {n}, 2), got shape {shape}.")
        s = s.astype(complex)
        if validate_norm:
            norms = np.sum(np.abs(s) ** 2, axis=-1)
            if not np.allclose(norms, 1.0, atol=self.ATOL):
                bad = np.argwhere(np.abs(norms - 1.0) > self.ATOL)
                raise ValueError(f"Per-qubit amplitudes must each have unit norm. Entries {bad.tolist()} do not.")
        self.data = [s]

    @classmethod
    def from_basis_state(cls, basis_state, wires=None) -> "ProductState":
        if isinstance(basis_state, BasisState):
            bits, wires = basis_state.parameters[0], basis_state.wires
        elif hasattr(basis_state, "parameters") and hasattr(basis_state, "wires") and wires is None:
            bits, wires = np.asarray(basis_state.parameters[0]).astype(int), basis_state.wires  # qml.BasisState
        else:
            assert wires is not None, "Must provide wires if basis_state is not a BasisState"
            bits = np.asarray(basis_state).astype(int)
        n = len(_wires_tuple(wires))
        if bits.ndim not in (1, 2) or bits.shape[-1] != n:
            raise ValueError(f"basis_state must have shape ({n},) or (batch, {n}) matching wires, got {bits.shape}.")
        return cls(np.eye(2, dtype=complex)[bits], wires=wires, validate_norm=False)

    @property
    def batch_size(self) -> Optional[int]:
        return None if self.data[0].ndim == 2 else self.data[0].shape[0]

    @property
    def is_basis_state(self):
        cached = getattr(self, "_is_basis_cache", None)  # the amplitudes of a state-prep record never change
        if cached is None:
            s = self.data[0]
            r = np.all((np.abs(s[..., 0]) < self.ATOL) | (np.abs(s[..., 1]) < self.ATOL), axis=-1)
            cached = self._is_basis_cache = bool(r) if self.batch_size is None else r
        return cached

    def basis_state_bits(self) -> np.ndarray:
        if not np.all(self.is_basis_state):
            raise ValueError("ProductState is not a computational-basis state.")
        bits = getattr(self, "_bits_cache", None)
        if bits is None:
            s = self.data[0]
            bits = self._bits_cache = (np.abs(s[..., 1]) > np.abs(s[..., 0])).astype(int)
        return bits

    def bloch(self):
        s = self.data[0]
        a, b = s[..., 0], s[..., 1]
        return 2.0 * np.real(np.conj(a) * b), 2.0 * np.imag(np.conj(a) * b), np.abs(a) ** 2 - np.abs(b) ** 2

    @property
    def covariance_matrix(self) -> np.ndarray:
        r"""``(\Lambda_0)_{\mu\nu} = i<c_\mu c_\nu>`` from the Bloch vectors (product_state.py:141-245).  Host-side
        INSPECTION property with the reference's name (O(n^2), numpy); the device never calls it -- its initial covariance
        comes from the ``mcb200_product_state_cov`` kernel (``ops.product_state_cov``), and this property is what the CPU
        tests compare that kernel's formula against."""
        x, y, z = self.bloch()
        n = x.shape[-1]
        cov = np.zeros(x.shape[:-1] + (2 * n, 2 * n))
        k = np.arange(n)
        cov[..., 2 * k, 2 * k + 1] = -z
        # parity strings p_jk = prod_{j<l<k} z_l via running products along each row j
        for j in range(n - 1):
            p = np.ones(x.shape[:-1])
            for kk in range(j + 1, n):
                cov[..., 2 * j, 2 * kk] = y[..., j] * p * x[..., kk]
                cov[..., 2 * j, 2 * kk + 1] = y[..., j] * p * y[..., kk]
                cov[..., 2 * j + 1, 2 * kk] = -x[..., j] * p * x[..., kk]
                cov[..., 2 * j + 1, 2 * kk + 1] = -x[..., j] * p * y[..., kk]
                p = p * z[..., kk]
        return cov - np.swapaxes(cov, -1, -2)

    @property
    def displacement_vector(self) -> np.ndarray:
        """``d[2k] = (prod_{l<k} z_l) x_k``, ``d[2k+1] = (prod_{l<k} z_l) y_k`` (_extended_covariance.py:12-65)."""
        x, y, z = self.bloch()
        zp = np.concatenate([np.ones(x.shape[:-1] + (1,)), np.cumprod(z[..., :-1], axis=-1)], axis=-1)
        d = np.zeros(x.shape[:-1] + (2 * x.shape[-1],))
        d[..., 0::2] = zp * x
        d[..., 1::2] = zp * y
        return d

    @property
    def extended_covariance_matrix(self) -> np.ndarray:
        """``[[cov, d], [-d^T, 0]]`` (_extended_covariance.py:68-98)."""
        cov, d = self.covariance_matrix, self.displacement_vector
        n2 = cov.shape[-1]
        out = np.zeros(cov.shape[:-2] + (n2 + 1, n2 + 1))
        out[..., :n2, :n2] = cov
        out[..., :n2, n2] = d
        out[..., n2, :n2] = -d
        return out
