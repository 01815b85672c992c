"""BASELINE.json's configurations at FULL size, checked through size-independent properties of the domain (the oracle
finishes only on small subsets at these sizes): probabilities lie in [0, 1] and sum to 1, SPTM columns are
orthonormal, Gram matrices are symmetric with a unit diagonal, identical samples have kernel value 1, two independent
CUDA routes agree, and a random subset agrees with the oracle.  Same synthetic inputs as bench.py."""
import numpy as np
import pytest
import torch

from oracle import covariance as cv
from oracle import gates, kernels, readout
from oracle.gates import Op
from tests.helpers import readme_ops

pytestmark = pytest.mark.gpu

from matchcake_b200 import ops as mops  # noqa: E402
from matchcake_b200.ml.kernels import FermionicPQCKernel  # noqa: E402

RTOL, ATOL = 1e-10, 1e-13


def _readme_gate_specs(n):
    g = []
    for w in range(0, n - 1, 2):
        for kind in (mops.GATE_RXRX, mops.GATE_RYRY, mops.GATE_RZRZ):
            g.append(mops.GateSpec(kind, w, w + 1, 0))
    for w in range(1, n - 1, 2):
        g.append(mops.GateSpec(mops.GATE_FSWAP, w, w + 1))
    return g


def test_c2_full_size_properties():
    """C2: N = 16 README circuit, projector |0..0>, 65 536 instances."""
    n, B = 16, 65536
    p_np = np.random.default_rng(0).uniform(0, 2 * np.pi, (B, 2))
    p = torch.as_tensor(p_np, device="cuda")
    circ = mops.CompiledCircuit(n, _readme_gate_specs(n), 2)
    fused = mops.circuit_expval_projector(circ, p)                       # one warp per instance, everything on chip
    x = mops.circuit_columns(circ, p)                                    # K1
    cov = mops.cov_basis(x, n)                                           # K2
    unfused = mops.probs(cov, torch.zeros((1, n), dtype=torch.int8, device="cuda"))[:, 0]   # K3
    np.testing.assert_allclose(fused.cpu().numpy(), unfused.cpu().numpy(), rtol=1e-11, atol=1e-15)
    assert float(fused.min()) >= 0.0 and float(fused.max()) <= 1.0 + 1e-12
    # R is orthogonal, Lambda is skew with Lambda Lambda^T = 1 (pure Gaussian state)
    eye = torch.eye(2 * n, dtype=torch.float64, device="cuda")
    assert float((x.transpose(1, 2) @ x - eye).abs().max()) < 1e-12
    assert float((cov + cov.transpose(1, 2)).abs().max()) == 0.0
    assert float((cov @ cov.transpose(1, 2) - eye).abs().max()) < 1e-12
    # the marginal over 4 wires sums to 1 for every instance
    cols = torch.arange(8, dtype=torch.int32, device="cuda")
    marg = mops.probs_all(mops.cov_basis(mops.circuit_columns(circ, p, cols), n), normalize=False)
    assert float((marg.sum(1) - 1.0).abs().max()) < 1e-12 and float(marg.min()) >= 0.0
    # a random subset against the oracle (lookup-table formulation of the reference)
    sel = np.random.default_rng(1).choice(B, 48, replace=False)
    zeros = np.zeros(n, dtype=int)
    ref = readout.probs_lookup_table(gates.chain(readme_ops(n, p_np[sel]), n), zeros, zeros, np.arange(n))
    np.testing.assert_allclose(fused.cpu().numpy()[sel], ref, rtol=RTOL, atol=ATOL)


def test_c3_full_size_properties():
    """C3: N = 32 FermionicPQCKernel Gram of 8192 samples (33.5 M evaluated cells)."""
    n, ns = 32, 8192
    x = np.random.default_rng(0).random((ns, 2 * n))
    x[ns - 1] = x[17]                                                     # identical samples -> kernel value 1
    k = FermionicPQCKernel(n_qubits=n, rotations="X,Z", r_dtype=torch.float64, c_dtype=torch.complex128)
    k.float32_gram = False
    k.fit(x)
    gram = k(torch.as_tensor(x).cuda())
    assert gram.shape == (ns, ns)
    assert float((gram - gram.T).abs().max()) == 0.0                      # mirrored, not recomputed
    assert float((torch.diagonal(gram) - 1.0).abs().max()) == 0.0
    assert float(gram.min()) >= 0.0 and float(gram.max()) <= 1.0 + 1e-12
    assert abs(float(gram[ns - 1, 17]) - 1.0) < 1e-12
    # a sub-Gram computed on its own equals the corresponding block (independent of the tile schedule)
    sub = k(torch.as_tensor(x[:96]).cuda())
    np.testing.assert_allclose(gram[:96, :96].cpu().numpy(), sub.cpu().numpy(), rtol=1e-12, atol=1e-15)
    # random cells against the reference's per-pair formulation (R = R_i R_j^T, lookup-table read-out)
    o = kernels.FermionicPQCKernelOracle(n_qubits=n, rotations="X,Z").fit(x)
    rng = np.random.default_rng(2)
    idx = np.stack([rng.integers(0, ns, 24), rng.integers(0, ns, 24)], axis=1)
    rows = np.unique(idx)
    sptm = {r: o.x_to_sptm(x[r:r + 1])[0] for r in rows}
    s0 = np.stack([sptm[i] for i in idx[:, 0]])
    s1 = np.stack([sptm[j] for j in idx[:, 1]])
    ref = o.pair_expvals(s0, s1, np.stack([np.arange(24), np.arange(24)], axis=1))
    got = gram.cpu().numpy()[idx[:, 0], idx[:, 1]]
    same = idx[:, 0] == idx[:, 1]
    np.testing.assert_allclose(got[~same], ref[~same], rtol=RTOL, atol=ATOL)


def test_c4_full_size_properties():
    """C4: N = 64, depth-64 brickwork, probabilities over 8 wires, 16 384 instances (528 MB of parameters)."""
    n, B, depth, kq = 64, 16384, 64, 8
    gate_list, names, off = [], [], 0
    kinds = (mops.GATE_RXRX, mops.GATE_RYRY, mops.GATE_RZRZ)
    knames = ("SptmCompRxRx", "SptmCompRyRy", "SptmCompRzRz")
    for layer in range(depth):
        for w in range(layer % 2, n - 1, 2):
            gate_list.append(mops.GateSpec(kinds[layer % 3], w, w + 1, off))
            names.append((knames[layer % 3], w, off))
            off += 2
    circ = mops.CompiledCircuit(n, gate_list, off)
    gen = torch.Generator(device="cuda").manual_seed(0)
    params = torch.rand((B, off), dtype=torch.float64, device="cuda", generator=gen) * (2 * np.pi)
    maj = torch.arange(2 * kq, dtype=torch.int32, device="cuda")
    xcols = mops.circuit_columns(circ, params, maj)
    eye = torch.eye(2 * kq, dtype=torch.float64, device="cuda")
    assert float((xcols.transpose(1, 2) @ xcols - eye).abs().max()) < 1e-11          # columns of an orthogonal matrix
    cov = mops.cov_basis(xcols, n)
    raw = mops.probs_all(cov, normalize=False)
    assert raw.shape == (B, 256) and float(raw.min()) >= 0.0
    assert float((raw.sum(1) - 1.0).abs().max()) < 1e-11                              # a probability distribution
    norm = mops.probs_all(cov, normalize=True)
    np.testing.assert_allclose(norm.cpu().numpy(), (raw / raw.sum(1, keepdim=True)).cpu().numpy(), rtol=1e-13)
    # the elimination tree against the per-outcome Pfaffians (two independent kernels) on a slice
    states = torch.as_tensor(readout.states_to_binary(np.arange(256), kq).copy(), dtype=torch.int8, device="cuda")
    np.testing.assert_allclose(raw[:512].cpu().numpy(), mops.probs(cov[:512], states).cpu().numpy(), rtol=1e-10, atol=1e-14)
    # three instances against the oracle (63 dense 128 x 128 products each)
    sel = [0, 7777, B - 1]
    pp = params[sel].cpu().numpy()
    ops_ = [Op(name, (w, w + 1), params=pp[:, o:o + 2]) for name, w, o in names]
    r_ref = gates.chain(ops_, n)
    ref = readout.analytic_probability(r_ref, cv.amplitudes_from_basis_state(np.zeros(n, int)), list(range(kq)))
    np.testing.assert_allclose(norm[sel].cpu().numpy(), ref, rtol=1e-10, atol=1e-13)


def test_c5_full_size_properties():
    """C5: N = 128 Nystroem Gram, 4096 landmarks x 262 144 samples (1.07 G evaluated cells, 8.6 GB of output)."""
    n, ns, nl = 128, 262144, 4096
    x = np.random.default_rng(0).random((ns, n))
    k = FermionicPQCKernel(n_qubits=n, rotations="X,Z", r_dtype=torch.float64, c_dtype=torch.complex128)
    k.float32_gram = False
    k.fit(x[:8])
    from sklearn.utils import check_random_state

    land_idx = check_random_state(0).permutation(ns)[:nl]
    xd, ld = torch.as_tensor(x).cuda(), torch.as_tensor(x[land_idx]).cuda()
    kmm = k(ld)
    assert float((kmm - kmm.T).abs().max()) == 0.0 and float((torch.diagonal(kmm) - 1.0).abs().max()) == 0.0
    knm = k(xd, ld, full_rectangle=True)                                   # every cell evaluated
    assert knm.shape == (ns, nl)
    assert float(knm.min()) >= 0.0 and float(knm.max()) <= 1.0 + 1e-12
    # a landmark against itself is 1; the landmark rows of the rectangle reproduce the square Gram
    rows = torch.as_tensor(land_idx[:512], device="cuda")
    cols = torch.arange(512, device="cuda")
    assert float((knm[rows, cols] - 1.0).abs().max()) < 1e-12
    blk = knm[rows][:, :512]
    off_diag = ~torch.eye(512, dtype=torch.bool, device="cuda")
    # (the mirrored half of kmm is cell (j, i): same number, the 64 factors rounded in the other operand order)
    np.testing.assert_allclose(blk[off_diag].cpu().numpy(), kmm[:512, :512][off_diag].cpu().numpy(), rtol=1e-10, atol=1e-300)
    # block-factorised route vs the dense 256 x 256 Pfaffian route, and vs the oracle, on a few cells
    k.use_block_factorisation = False
    dense = k(xd[:6], ld[:5], full_rectangle=True)
    np.testing.assert_allclose(knm[:6, :5].cpu().numpy(), dense.cpu().numpy(), rtol=1e-10, atol=1e-300)
    o = kernels.FermionicPQCKernelOracle(n_qubits=n, rotations="X,Z").fit(x[:8])
    s0, s1 = o.x_to_sptm(x[:3]), o.x_to_sptm(x[land_idx[:2]])
    idx = np.array([[0, 0], [1, 1], [2, 0]])
    ref = o.pair_expvals(s0, s1, idx)
    # the reference reads out sqrt(|det| + 1e-32) (utils/_pfaffian.py:110-118); at N = 128 the kernel values are small
    # enough (1e-13 .. 1e-15) for that floor to show: it is applied inside the Gram kernels (kernel.pfaffian_epsilon)
    assert k.pfaffian_epsilon == 1e-32
    got = knm[:3, :2].cpu().numpy()[idx[:, 0], idx[:, 1]]
    np.testing.assert_allclose(got, ref, rtol=1e-10, atol=1e-300)
    k.pfaffian_epsilon = 0.0
    exact = k(xd[:3], ld[:2], full_rectangle=True).cpu().numpy()[idx[:, 0], idx[:, 1]]
    np.testing.assert_allclose(np.sqrt(exact ** 2 + 1e-32), got, rtol=1e-13, atol=0)
    assert np.max(np.abs(exact / got - 1.0)) > 1e-7   # the floor is not a no-op at these magnitudes
