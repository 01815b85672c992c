"""Gate -> single-particle transition matrix (SPTM) and the SPTM chain.  TEST INFRASTRUCTURE.

numpy float64/complex128 restatement of (paths relative to /root/reference/src/matchcake):

* closed-form SPTM blocks ....... operations/single_particle_transition_matrices/
                                  sptm_comp_rxrx.py:48-57, sptm_comp_ryry.py:53-67,
                                  sptm_comp_rzrz.py:51-67, sptm_fswap.py:12-27
* matchgate 4x4 from rotations .. operations/comp_rotations.py:18-80
* generic gate -> SPTM .......... utils/__init__.py:561-596 (1/4 Tr(U c_mu U^dag c_nu)),
                                  utils/majorana.py:15-47 (c_2k = Z..ZX, c_2k+1 = Z..ZY)
* padding ....................... single_particle_transition_matrix.py:379-402
* block-diagonal assembly ....... single_particle_transition_matrix.py:228-271
* "neighbours" contraction ...... devices/contraction_strategies/neighbours_strategy.py:13-26,
                                  contraction_container.py:95-118, device_utils.py:18-56
* product order ................. constants/__init__.py:24,33; utils/math.py:370-444;
                                  devices/nif_device.py:168-195 (R <- R_old @ R_new)
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np

PAULI_I = np.eye(2, dtype=complex)
PAULI_X = np.array([[0, 1], [1, 0]], dtype=complex)
PAULI_Y = np.array([[0, -1j], [1j, 0]], dtype=complex)
PAULI_Z = np.array([[1, 0], [0, -1]], dtype=complex)

REAL_PART_ATOL = 1e-6  # single_particle_transition_matrix.py:69


# --------------------------------------------------------------------------- closed forms
def sptm_comp_rxrx(params: np.ndarray) -> np.ndarray:
    """sptm_comp_rxrx.py:48-57."""
    params = np.asarray(params, dtype=np.float64)
    m = np.zeros(params.shape[:-1] + (4, 4), dtype=np.float64)
    theta, phi = params[..., 0] / 2, params[..., 1] / 2
    m[..., 0, 0] = np.cos(phi - theta)
    m[..., 0, 3] = -np.sin(phi - theta)
    m[..., 1, 1] = np.cos(phi + theta)
    m[..., 1, 2] = np.sin(phi + theta)
    m[..., 2, 1] = -np.sin(phi + theta)
    m[..., 2, 2] = np.cos(phi + theta)
    m[..., 3, 0] = np.sin(phi - theta)
    m[..., 3, 3] = np.cos(phi - theta)
    return m


def sptm_comp_ryry(params: np.ndarray) -> np.ndarray:
    """sptm_comp_ryry.py:53-67 (angle clipping is off by default, :23)."""
    params = np.asarray(params, dtype=np.float64)
    m = np.zeros(params.shape[:-1] + (4, 4), dtype=np.float64)
    theta, phi = params[..., 0], params[..., 1]
    tp = (theta + phi) / 2
    tm = (theta - phi) / 2
    m[..., 0, 0] = np.cos(tp)
    m[..., 0, 2] = -np.sin(tp)
    m[..., 1, 1] = np.cos(tm)
    m[..., 3, 1] = -np.sin(tm)
    m[..., 2, 0] = np.sin(tp)
    m[..., 2, 2] = np.cos(tp)
    m[..., 1, 3] = np.sin(tm)
    m[..., 3, 3] = np.cos(tm)
    return m


def sptm_comp_rzrz(params: np.ndarray) -> np.ndarray:
    """sptm_comp_rzrz.py:51-67: built complex, imaginary part (~1e-16) dropped by
    ``_enforce_real`` (single_particle_transition_matrix.py:87-113)."""
    params = np.asarray(params, dtype=np.float64)
    m = np.zeros(params.shape[:-1] + (4, 4), dtype=complex)
    theta, phi = params[..., 0], params[..., 1]
    exp_theta, exp_phi = np.exp(1j * theta), np.exp(1j * phi)
    h = (theta + phi) / 2
    exp_tp = np.exp(-1j * h)
    m[..., 0, 0] = np.cos(h)
    m[..., 0, 1] = np.sin(h)
    m[..., 1, 0] = -np.sin(h)
    m[..., 1, 1] = np.cos(h)
    m[..., 2, 2] = (exp_theta + exp_phi) * exp_tp / 2
    m[..., 2, 3] = 1j * (exp_phi - exp_theta) * exp_tp / 2
    m[..., 3, 2] = 1j * (exp_theta - exp_phi) * exp_tp / 2
    m[..., 3, 3] = (exp_theta + exp_phi) * exp_tp / 2
    return enforce_real(m)


def sptm_fswap(wires: Sequence[int]) -> np.ndarray:
    """sptm_fswap.py:12-27 -- signed-permutation-free cyclic block, non-adjacent wires allowed."""
    wires_arr = np.asarray(wires)
    wire0, wire1 = np.sort(wires_arr)
    size = 2 * (wire1 - wire0 + 1)
    m = np.zeros((size, size), dtype=np.float64)
    first_idx = 2 * (wire1 - wire0)
    m[:2, first_idx:first_idx + 2] = np.eye(2)
    rows, cols = np.where(np.eye(size, k=-2))
    m[rows, cols] = 1
    if wire0 != wires_arr[0]:
        m = m.T
    return m


def enforce_real(matrix: np.ndarray, check_real: bool = True) -> np.ndarray:
    """single_particle_transition_matrix.py:87-113."""
    if not np.iscomplexobj(matrix):
        return matrix
    if check_real:
        max_abs_imag = np.max(np.abs(matrix.imag)) if matrix.size else 0.0
        if max_abs_imag > REAL_PART_ATOL:
            raise ValueError(
                f"The single-particle transition matrix must be real, but its maximum absolute "
                f"imaginary part is {max_abs_imag}, which exceeds {REAL_PART_ATOL}."
            )
    return np.ascontiguousarray(matrix.real)


# --------------------------------------------------------------------------- matchgates (4x4 complex)
def _rot_matrix(param: np.ndarray, direction: str) -> np.ndarray:
    """comp_rotations.py:18-44."""
    param = np.asarray(param, dtype=np.float64)
    m = np.zeros(param.shape + (2, 2), dtype=complex)
    if direction == "X":
        m[..., 0, 0] = np.cos(param / 2)
        m[..., 0, 1] = -1j * np.sin(param / 2)
        m[..., 1, 0] = -1j * np.sin(param / 2)
        m[..., 1, 1] = np.cos(param / 2)
    elif direction == "Y":
        m[..., 0, 0] = np.cos(param / 2)
        m[..., 0, 1] = -np.sin(param / 2)
        m[..., 1, 0] = np.sin(param / 2)
        m[..., 1, 1] = np.cos(param / 2)
    elif direction == "Z":
        m[..., 0, 0] = np.exp(-1j * param / 2)
        m[..., 1, 1] = np.exp(1j * param / 2)
    else:
        raise ValueError(f"Invalid direction {direction}.")
    return m


def matchgate_from_sub_matrices(outer: np.ndarray, inner: np.ndarray) -> np.ndarray:
    """matchgate_standard_params.py:96-104: M = [[a,0,0,b],[0,w,x,0],[0,y,z,0],[c,0,0,d]]."""
    outer = np.asarray(outer, dtype=complex)
    inner = np.asarray(inner, dtype=complex)
    shape = np.broadcast_shapes(outer.shape[:-2], inner.shape[:-2])
    m = np.zeros(shape + (4, 4), dtype=complex)
    m[..., 0, 0] = outer[..., 0, 0]
    m[..., 0, 3] = outer[..., 0, 1]
    m[..., 3, 0] = outer[..., 1, 0]
    m[..., 3, 3] = outer[..., 1, 1]
    m[..., 1, 1] = inner[..., 0, 0]
    m[..., 1, 2] = inner[..., 0, 1]
    m[..., 2, 1] = inner[..., 1, 0]
    m[..., 2, 2] = inner[..., 1, 1]
    return m


def matchgate_comp_rotation(params: np.ndarray, directions: Sequence[str]) -> np.ndarray:
    """comp_rotations.py:47-80: outer = R_{dir[1]}(params[...,0]), inner = R_{dir[0]}(params[...,1])."""
    params = np.asarray(params, dtype=np.float64)
    outer = _rot_matrix(params[..., 0], directions[1])
    inner = _rot_matrix(params[..., 1], directions[0])
    return matchgate_from_sub_matrices(outer, inner)


def matchgate_fswap() -> np.ndarray:
    """fermionic_swap.py:15-52: M(Z, X)."""
    return matchgate_from_sub_matrices(PAULI_Z, PAULI_X)


def majorana(i: int, n: int) -> np.ndarray:
    """utils/majorana.py:15-47: even index -> X, odd index -> Y, Z-string on qubits < k."""
    k = i // 2
    gate = PAULI_Y if (i + 1) % 2 == 0 else PAULI_X
    out = np.eye(1, dtype=complex)
    for p in [PAULI_Z] * k + [gate] + [PAULI_I] * (n - k - 1):
        out = np.kron(out, p)
    return out


def sptm_from_gate(u: np.ndarray) -> np.ndarray:
    """utils/__init__.py:561-596: R_{mu nu} = 2^-n Tr(U c_mu U^dag c_nu); (..., 2^n, 2^n) -> (..., 2n, 2n) complex."""
    u = np.asarray(u, dtype=complex)
    n = int(np.log2(u.shape[-1]))
    c = np.stack([majorana(i, n) for i in range(2 * n)])  # (2n, 2^n, 2^n)
    u_c = np.einsum("...ij,mjq->...miq", u, c)
    u_c_ud = np.einsum("...miq,...kq->...mik", u_c, np.conjugate(u))
    u_c_ud_c = np.einsum("...kij,mjq->...kmiq", u_c_ud, c)
    traced = np.einsum("...ii", u_c_ud_c)
    return traced / c.shape[-1]


# --------------------------------------------------------------------------- op model
MATCHGATE_KINDS = {"CompRxRx": "XX", "CompRyRy": "YY", "CompRzRz": "ZZ"}
SPTM_CLOSED_FORMS = {"SptmCompRxRx": sptm_comp_rxrx, "SptmCompRyRy": sptm_comp_ryry, "SptmCompRzRz": sptm_comp_rzrz}
_MG_TO_SPTM = {"CompRxRx": "SptmCompRxRx", "CompRyRy": "SptmCompRyRy", "CompRzRz": "SptmCompRzRz"}


def make_wires_continuous(wires: Sequence[int]) -> Tuple[int, ...]:
    return tuple(range(min(wires), max(wires) + 1))


@dataclass
class Op:
    """A gate as the device sees it.

    ``kind``:
      * matchgate flavour (``MatchgateOperation`` subclasses): ``CompRxRx``, ``CompRyRy``,
        ``CompRzRz`` (``params``), ``fSWAP`` (no data), ``Matchgate`` (``matrix`` = (...,4,4) complex)
      * SPTM flavour (``SingleParticleTransitionMatrixOperation`` subclasses): ``SptmCompRxRx``,
        ``SptmCompRyRy``, ``SptmCompRzRz`` (``params``), ``SptmFSwap``, ``Sptm`` (``matrix`` real (...,2w,2w))
    """

    kind: str
    wires: Tuple[int, ...]
    params: Optional[np.ndarray] = None
    matrix: Optional[np.ndarray] = None

    @property
    def is_matchgate(self) -> bool:
        return self.kind in ("CompRxRx", "CompRyRy", "CompRzRz", "fSWAP", "Matchgate")

    @property
    def cs_wires(self) -> Tuple[int, ...]:
        return make_wires_continuous(self.wires)

    @property
    def sorted_wires(self) -> Tuple[int, ...]:
        return tuple(sorted(self.wires))

    def matchgate_matrix(self) -> np.ndarray:
        if self.kind in MATCHGATE_KINDS:
            return matchgate_comp_rotation(self.params, MATCHGATE_KINDS[self.kind])
        if self.kind == "fSWAP":
            return matchgate_fswap()
        if self.kind == "Matchgate":
            return np.asarray(self.matrix, dtype=complex)
        raise ValueError(self.kind)

    def to_sptm(self) -> "Op":
        """``SingleParticleTransitionMatrixOperation.from_operation`` (:184-226)."""
        if not self.is_matchgate:
            return self
        if self.kind in _MG_TO_SPTM:  # comp_rotations.py:247-285 -> closed forms
            return Op(_MG_TO_SPTM[self.kind], self.wires, params=self.params)
        if self.kind == "fSWAP":  # fermionic_swap.py:50-51
            return Op("SptmFSwap", self.wires)
        m = enforce_real(sptm_from_gate(self.matchgate_matrix()))  # matchgate_operation.py:342,434-437
        return Op("Sptm", self.wires, matrix=m)

    def sptm_matrix(self) -> np.ndarray:
        """Real (..., 2w, 2w) block on ``cs_wires``."""
        if self.is_matchgate:
            return self.to_sptm().sptm_matrix()
        if self.kind in SPTM_CLOSED_FORMS:
            return SPTM_CLOSED_FORMS[self.kind](self.params)
        if self.kind == "SptmFSwap":
            return sptm_fswap(self.wires)
        if self.kind == "Sptm":
            return enforce_real(np.asarray(self.matrix))
        raise ValueError(self.kind)

    def adjoint(self) -> "Op":
        """single_particle_transition_matrix.py:420-421 (transpose); fSWAP is self-adjoint (sptm_fswap.py:29)."""
        if self.kind == "SptmFSwap":
            return self
        return Op("Sptm", self.wires, matrix=np.swapaxes(self.sptm_matrix(), -1, -2))


def pad(matrix: np.ndarray, op_cs_wires: Sequence[int], wires: Sequence[int]) -> np.ndarray:
    """single_particle_transition_matrix.py:379-402: embed the block at offset 2*index(wire0) of I."""
    cs_wires = make_wires_continuous(wires)
    if tuple(op_cs_wires) == cs_wires:
        return matrix
    d = 2 * len(cs_wires)
    out = np.zeros(matrix.shape[:-2] + (d, d), dtype=matrix.dtype)
    out[..., :, :] = np.eye(d)
    w0 = cs_wires.index(op_cs_wires[0])
    out[..., 2 * w0:2 * w0 + matrix.shape[-2], 2 * w0:2 * w0 + matrix.shape[-1]] = matrix
    return out


def sptm_matmul(first: Op, second: Op) -> Op:
    """``Sptm.__matmul__`` (:404-418) with ``_FOP_MATMUL_DIRECTION = RL`` (first @ second)."""
    all_wires = make_wires_continuous(sorted(set(first.wires) | set(second.wires)))
    a = pad(first.sptm_matrix(), first.cs_wires, all_wires)
    b = pad(second.sptm_matrix(), second.cs_wires, all_wires)
    return Op("Sptm", all_wires, matrix=np.einsum("...ij,...jk->...ik", a, b))


def circuit_or_fop_matmul(first: Op, second: Op) -> Op:
    """device_utils.py:18-56."""
    if first.is_matchgate and second.is_matchgate:
        # circuit_matmul with LR direction => second @ first (matchgate_operation.py:319-340)
        if first.wires != second.wires:
            return sptm_matmul(second.to_sptm(), first.to_sptm())
        u = np.einsum("...ij,...jk->...ik", second.matchgate_matrix(), first.matchgate_matrix())
        return Op("Matchgate", first.wires, matrix=u)
    return sptm_matmul(first.to_sptm(), second.to_sptm())


def from_operations(ops: List[Op]) -> Optional[Op]:
    """single_particle_transition_matrix.py:228-271 -- block-diagonal assembly on the continuous span."""
    if len(ops) == 0:
        return None
    ops = [op.to_sptm() for op in ops]
    if len(ops) == 1:
        return ops[0]
    all_wires = make_wires_continuous(sorted(set(w for op in ops for w in op.cs_wires)))
    mats = [op.sptm_matrix() for op in ops]
    batch_sizes = [m.shape[0] for m in mats if m.ndim == 3]
    d = 2 * len(all_wires)
    if batch_sizes:
        matrix = np.zeros((batch_sizes[0], d, d), dtype=complex)
        matrix[:, ...] = np.eye(d)
    else:
        matrix = np.eye(d, dtype=complex)
    for op, m in zip(ops, mats):
        w0 = all_wires.index(op.sorted_wires[0])
        matrix[..., 2 * w0:2 * w0 + m.shape[-2], 2 * w0:2 * w0 + m.shape[-1]] = m
    return Op("Sptm", all_wires, matrix=enforce_real(matrix))


class NeighboursContainer:
    """neighbours_strategy.py:13-26 + contraction_container.py:95-118."""

    def __init__(self):
        self.op_container = {}
        self.wires_set = set()

    def try_add(self, op: Op) -> bool:
        wires = op.cs_wires
        if wires in self.op_container:
            self.op_container[wires] = circuit_or_fop_matmul(self.op_container[wires], op)
            return True
        if any(w in self.wires_set for w in wires):
            return False
        self.op_container[wires] = op
        self.wires_set.update(wires)
        return True

    def contract_and_clear(self) -> Optional[Op]:
        out = from_operations(list(self.op_container.values()))
        self.op_container.clear()
        self.wires_set.clear()
        return out

    def push_contract(self, op: Op) -> Optional[Op]:
        if not self.try_add(op):
            contracted = self.contract_and_clear()
            assert self.try_add(op)
            return contracted
        return None


def chain(ops: Iterable[Op], n_wires: int, contraction: Optional[str] = "neighbours") -> np.ndarray:
    """``apply_generator`` (nif_device.py:281-347) + ``apply_op`` (:381-400):
    R <- R @ pad(R_op), einsum("...ij,...jk->...ik") (:195).  Returns (d,d) or (B,d,d)."""
    wires = tuple(range(n_wires))
    state = {"R": None}

    def apply_op(op: Op):
        m = pad(op.to_sptm().sptm_matrix(), op.cs_wires, wires).astype(np.float64)
        state["R"] = m if state["R"] is None else np.einsum("...ij,...jk->...ik", state["R"], m)

    if contraction in (None, "none"):
        for op in ops:
            apply_op(op)
    elif contraction == "neighbours":
        container = NeighboursContainer()
        for op in ops:
            out = container.push_contract(op)
            if out is not None:
                apply_op(out)
        last = container.contract_and_clear()
        if last is not None:
            apply_op(last)
    else:
        raise ValueError(f"Unknown contraction strategy {contraction}")
    if state["R"] is None:  # nif_device.py:1174-1179
        return np.eye(2 * n_wires)
    return state["R"]


def random_sptm(n_wires: int, seed=None, batch_size=None) -> np.ndarray:
    """``SingleParticleTransitionMatrixOperation.random_params`` (:274-287):
    default_rng(seed).normal upper triangle -> expm(4 h)."""
    from scipy.linalg import expm

    rng = np.random.default_rng(seed)
    d = 2 * n_wires
    iu = np.triu_indices(d, k=1)
    shape = ([batch_size] if batch_size is not None else [])
    rn = rng.normal(size=shape + [iu[0].size])
    m = np.zeros(shape + [d, d], dtype=complex)
    m[..., iu[0], iu[1]] = rn
    m[..., iu[1], iu[0]] = -rn
    return enforce_real(expm(4 * m))
