"""torch-CPU restatement of the reference's op sequence, used ONLY as the timed CPU baseline.  TEST INFRASTRUCTURE.

The reference (MatchCake v0.2.3) cannot be installed in this image (PennyLane / torch_pfaffian / autoray are
absent, no network), so ``bench.py --impl reference`` and the ``cpu_baseline`` leg time this module instead: the
same sequence of torch float64/complex128 operations the reference issues on its CPU path, with all host
threads, but WITHOUT its per-gate Python object construction / matchgate validation and WITHOUT the
``np.ndindex`` pair generator -- i.e. an optimistic (fast) stand-in ("B-ref").

Op sequence followed (paths relative to /root/reference/src/matchcake):
  a1/a2  operations/comp_rotations.py:47-80, matchgate_operation.py:319-340   (4x4 complex matchgates, fused per pair)
  a3     utils/__init__.py:561-596                                            (matchgate -> SPTM by four einsums)
  a5/a6  single_particle_transition_matrix.py:228-271, 379-402                (block-diagonal layer, dense pad)
  a7     devices/nif_device.py:168-195                                        (dense einsum chain)
  a8     utils/__init__.py:352-369                                            (transition matrix T)
  a9     base/lookup_table.py:132-205, 619-710                                (complex lookup-table observable)
  a13    utils/_pfaffian.py:110-118                                           (sqrt|det| via torch.linalg.det)
  a16    ml/kernels/nif_kernel.py:159-221                                     (per-pair R = b0 @ b1^T + read-out)
Every function is checked against the numpy oracle in tests/test_oracle_golden.py.
"""
from __future__ import annotations

import numpy as np
import torch

from . import gates as og

C = torch.complex128
R = torch.float64


def _majorana_tensor(n: int) -> torch.Tensor:
    return torch.as_tensor(np.stack([og.majorana(i, n) for i in range(2 * n)]), dtype=C)


_MAJ2 = None


def _rot(param: torch.Tensor, direction: str) -> torch.Tensor:
    m = torch.zeros(param.shape + (2, 2), dtype=C)
    if direction == "X":
        m[..., 0, 0] = torch.cos(param / 2)
        m[..., 0, 1] = -1j * torch.sin(param / 2)
        m[..., 1, 0] = -1j * torch.sin(param / 2)
        m[..., 1, 1] = torch.cos(param / 2)
    elif direction == "Y":
        m[..., 0, 0] = torch.cos(param / 2)
        m[..., 0, 1] = -torch.sin(param / 2)
        m[..., 1, 0] = torch.sin(param / 2)
        m[..., 1, 1] = torch.cos(param / 2)
    else:
        m[..., 0, 0] = torch.exp(-1j * param / 2)
        m[..., 1, 1] = torch.exp(1j * param / 2)
    return m


def _matchgate(params: torch.Tensor, direction: str) -> torch.Tensor:
    outer, inner = _rot(params[..., 0], direction), _rot(params[..., 1], direction)
    m = torch.zeros(params.shape[:-1] + (4, 4), dtype=C)
    m[..., 0, 0], m[..., 0, 3], m[..., 3, 0], m[..., 3, 3] = outer[..., 0, 0], outer[..., 0, 1], outer[..., 1, 0], outer[..., 1, 1]
    m[..., 1, 1], m[..., 1, 2], m[..., 2, 1], m[..., 2, 2] = inner[..., 0, 0], inner[..., 0, 1], inner[..., 1, 0], inner[..., 1, 1]
    return m


def _sptm_from_gate(u: torch.Tensor) -> torch.Tensor:
    global _MAJ2
    if _MAJ2 is None:
        _MAJ2 = _majorana_tensor(2)
    c = _MAJ2
    u_c = torch.einsum("...ij,mjq->...miq", u, c)
    u_c_ud = torch.einsum("...miq,...kq->...mik", u_c, torch.conj(u))
    u_c_ud_c = torch.einsum("...kij,mjq->...kmiq", u_c_ud, c)
    return torch.einsum("...ii", u_c_ud_c) / 4


def readme_circuit_sptm(params: torch.Tensor, n: int) -> torch.Tensor:
    """README.md:48-66 circuit -> global SPTM (B, d, d) the way the device builds it (neighbours contraction)."""
    B = params.shape[0]
    d = 2 * n
    u = _matchgate(params, "Z") @ _matchgate(params, "Y") @ _matchgate(params, "X")  # second @ first, fused per pair
    blk = torch.real(_sptm_from_gate(u))
    layer1 = torch.zeros((B, d, d), dtype=C)
    layer1[:, ...] = torch.eye(d, dtype=C)
    for w in range(0, n - 1, 2):
        layer1[:, 2 * w:2 * w + 4, 2 * w:2 * w + 4] = blk.to(C)
    r = torch.real(layer1)
    fs = torch.as_tensor(og.sptm_fswap((0, 1)), dtype=R)
    odd = list(range(1, n - 1, 2))
    if odd:
        span = 2 * (odd[-1] + 1 - odd[0] + 1)
        layer2 = torch.eye(span, dtype=C)
        for w in odd:
            o = 2 * (w - odd[0])
            layer2[o:o + 4, o:o + 4] = fs.to(C)
        padded = torch.eye(d, dtype=R)
        padded[2 * odd[0]:2 * odd[0] + span, 2 * odd[0]:2 * odd[0] + span] = torch.real(layer2)
        r = torch.einsum("...ij,...jk->...ik", r, padded)
    return r


def projector_zero_lookup(r: torch.Tensor) -> torch.Tensor:
    """<0..0| rho |0..0> for basis input |0..0> through the complex lookup table (a8-a10, a13).  r: (B, d, d)."""
    d = r.shape[-1]
    n = d // 2
    rt = torch.einsum("...ij->...ji", r)
    t = 0.5 * (rt[..., ::2, :] + 1j * rt[..., 1::2, :])
    bm = torch.as_tensor(og.PAULI_I[:0], dtype=C)  # placeholder to keep dtype import local
    bm = torch.zeros((d, d), dtype=C)
    for i in range(n):
        bm[2 * i:2 * i + 2, 2 * i:2 * i + 2] = torch.tensor([[1, 1j], [-1j, 1]], dtype=C)
    tc = torch.conj(t)
    tb = torch.einsum("...ij,jk->...ik", t, bm)
    btt = torch.einsum("ij,...kj->...ik", bm, t)
    btd = torch.einsum("ij,...kj->...ik", bm, tc)
    items = torch.stack([
        torch.einsum("...pj,...kj->...pk", tb, t), torch.einsum("...pj,...kj->...pk", tb, tc),
        torch.einsum("...pi,...ik->...pk", tc, btt), torch.einsum("...pi,...ik->...pk", tc, btd)])  # (4, B, n, n)
    size = 2 * n
    iu = np.triu_indices(size, k=1)
    cls = np.tile(np.array([0, 1]), n)            # target bit 0 -> classes (0, 1) per wire
    pos = np.repeat(np.arange(n), 2)
    which = torch.as_tensor(2 * cls[iu[0]] + cls[iu[1]])
    vals = items[which, :, torch.as_tensor(pos[iu[0]]), torch.as_tensor(pos[iu[1]])]  # (P, B)
    obs = torch.zeros((r.shape[0], size, size), dtype=C)
    obs[:, iu[0], iu[1]] = vals.transpose(0, 1)
    obs = obs - torch.einsum("...ij->...ji", obs)
    if size > 128:
        # one matrix per LAPACK call: the batched complex LU of this torch/MKL build deadlocks for 256 x 256 matrices once
        # torch.set_num_threads() has been called (nested threading); a loop is also what the reference's batch of
        # one-pair circuits amounts to at this size
        det = torch.stack([torch.linalg.det(o) for o in obs])
    else:
        det = torch.linalg.det(obs)
    return torch.sqrt(torch.abs(det) + 1e-32)


def readme_expvals(params: np.ndarray, n: int) -> np.ndarray:
    """Config C2 on the CPU: README circuit, projector |0..0><0..0| on all wires."""
    p = torch.as_tensor(params, dtype=R)
    return projector_zero_lookup(readme_circuit_sptm(p, n)).numpy()


def gram_pairs(sptm0: np.ndarray, sptm1: np.ndarray, indices: np.ndarray) -> np.ndarray:
    """nif_kernel.py:210-221 for one batch of index pairs."""
    b0 = torch.as_tensor(sptm0[indices[:, 0]], dtype=R)
    b1 = torch.as_tensor(sptm1[indices[:, 1]], dtype=R)
    r = torch.einsum("...ij,...jk->...ik", b0, torch.einsum("...ij->...ji", b1))
    return projector_zero_lookup(r).numpy()
