"""GPU parity: the CUDA path, called through the C ABI (ctypes -> libmcb200.so), against the CPU oracle.

Tolerance: north_star asks for 1e-10 relative on expectation values and Gram entries.  Probabilities of
parity-forbidden outcomes are exactly 0 in exact arithmetic, where a relative bound is meaningless (SURVEY.md
section 7 "hard parts"), so every comparison uses rtol=1e-10 with atol=1e-13."""
import json
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import covariance as cv
from oracle import gates, kernels, pfaffian, readout
from oracle.gates import Op
from tests.helpers import ops_to_circuit, random_skew, readme_ops

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-10, 1e-13
G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_golden.json")))


def dev(x, dtype=torch.float64):
    return torch.as_tensor(np.ascontiguousarray(x)).to(device="cuda", dtype=dtype)


def compiled(ops, n, batch, **kw):
    from matchcake_b200 import ops as mops

    g, params, consts = ops_to_circuit(ops, n, batch)
    return mops.CompiledCircuit(n, g, params.shape[1], consts=consts, **kw), dev(params)


# ------------------------------------------------------------------ K3
@pytest.mark.parametrize("m", [2, 4, 6, 8, 16, 30, 32, 34, 62, 64, 66, 96, 128, 130, 136, 146, 200, 250, 256])
@pytest.mark.parametrize("signed", [True, False])
def test_pfaffian_matches_oracle(m, signed):
    from matchcake_b200 import ops as mops

    rng = np.random.default_rng(m)
    n = 37 if m <= 64 else 5
    a = random_skew(rng, (n, m, m)) / np.sqrt(m)
    ref = pfaffian.pfaffian_signed(a)
    if not signed:
        ref = np.abs(ref)
    out = mops.pfaffian(dev(a), signed=signed).cpu().numpy()
    np.testing.assert_allclose(out, ref, rtol=RTOL, atol=ATOL * np.max(np.abs(ref)))


def test_pfaffian_kats_and_edge_cases():
    from matchcake_b200 import ops as mops

    k = G["pfaffian_kats"]
    assert mops.pfaffian(dev(k["two_by_two"]["matrix"]), signed=True).item() == pytest.approx(3.0, abs=1e-14)
    a, b, c, d, e, f = k["four_by_four"]["abcdef"]
    m4 = np.array([[0, a, b, c], [-a, 0, d, e], [-b, -d, 0, f], [-c, -e, -f, 0]])
    assert mops.pfaffian(dev(m4), signed=True).item() == pytest.approx(8.0, abs=1e-13)
    pm = np.array(k["pivot_swap"]["matrix"])
    assert mops.pfaffian(dev(pm), signed=True).item() ** 2 == pytest.approx(np.linalg.det(pm), abs=1e-12)
    assert mops.pfaffian(dev(np.zeros((3, 0, 0))), signed=True).cpu().tolist() == [1.0, 1.0, 1.0]  # empty -> 1
    assert mops.pfaffian(dev(np.ones((2, 3, 3))), signed=True).cpu().tolist() == [0.0, 0.0]  # odd -> 0
    assert mops.pfaffian(dev(np.zeros((2, 4, 4))), signed=False).cpu().tolist() == [0.0, 0.0]  # singular -> 0
    assert mops.pfaffian(dev(np.zeros((0, 4, 4)))).shape == (0,)
    # leading batch axes and scale
    rng = np.random.default_rng(0)
    x = random_skew(rng, (2, 3, 8, 8))
    out = mops.pfaffian(dev(x), signed=True, scale=0.25).cpu().numpy()
    np.testing.assert_allclose(out, 0.25 * pfaffian.pfaffian_signed(x), rtol=RTOL)


def test_pfaffian_needs_pivoting():
    """Zero leading entries force row/column swaps at every step."""
    from matchcake_b200 import ops as mops

    rng = np.random.default_rng(3)
    m = 16
    a = random_skew(rng, (9, m, m))
    a[:, 0, 1] = a[:, 1, 0] = 0.0
    a[:, 2, 3] = a[:, 3, 2] = 0.0
    a[:, 0, 2:9] = 0.0
    a[:, 2:9, 0] = 0.0
    ref = pfaffian.pfaffian_signed(a)
    np.testing.assert_allclose(mops.pfaffian(dev(a), signed=True).cpu().numpy(), ref, rtol=RTOL, atol=1e-12)
    np.testing.assert_allclose(ref ** 2, np.linalg.det(a), rtol=1e-8)


# ------------------------------------------------------------------ K1
@pytest.mark.parametrize("n", [2, 4, 5, 16])
@pytest.mark.parametrize("matchgate", [True, False])
def test_circuit_sptm_readme(n, matchgate):
    from matchcake_b200 import ops as mops

    rng = np.random.default_rng(n)
    B = 7
    p = rng.uniform(0, 2 * np.pi, (B, 2))
    ops = readme_ops(n, p, matchgate)
    ref = np.broadcast_to(gates.chain(ops, n), (B, 2 * n, 2 * n))
    circ, params = compiled(ops, n, B)
    out = mops.circuit_columns(circ, params).cpu().numpy()
    np.testing.assert_allclose(out, ref, rtol=RTOL, atol=ATOL)
    cols = torch.tensor([3, 0, 2 * n - 1], dtype=torch.int32)
    sub = mops.circuit_columns(circ, params, cols).cpu().numpy()
    np.testing.assert_allclose(sub, ref[:, :, [3, 0, 2 * n - 1]], rtol=RTOL, atol=ATOL)


def test_circuit_sptm_general_gates():
    """Long-range fSWAPs in both orientations, explicit 4x4 blocks (shared and per-instance), empty circuit."""
    from matchcake_b200 import ops as mops

    rng = np.random.default_rng(11)
    n, B = 6, 5
    blk = gates.random_sptm(2, seed=1)
    blk_b = gates.random_sptm(2, seed=2, batch_size=B)
    ops = [Op("SptmCompRyRy", (1, 2), params=rng.uniform(0, 6, (B, 2))), Op("SptmFSwap", (0, 4)),
           Op("Sptm", (2, 3), matrix=blk), Op("SptmFSwap", (5, 2)), Op("SptmCompRzRz", (4, 5), params=rng.uniform(0, 6, (B, 2))),
           Op("Sptm", (0, 1), matrix=blk_b), Op("SptmFSwap", (3, 2)), Op("SptmCompRxRx", (3, 4), params=rng.uniform(0, 6, (B, 2)))]
    ref = gates.chain(ops, n, contraction=None)
    circ, params = compiled(ops, n, B)
    np.testing.assert_allclose(mops.circuit_columns(circ, params).cpu().numpy(), ref, rtol=RTOL, atol=ATOL)
    assert np.allclose(gates.chain(ops, n), ref, atol=1e-13)  # neighbours contraction == plain product
    empty = mops.CompiledCircuit(3, [], 0)
    out = mops.circuit_columns(empty, torch.zeros((2, 0), dtype=torch.float64, device="cuda")).cpu().numpy()
    np.testing.assert_array_equal(out, np.broadcast_to(np.eye(6), (2, 6, 6)))


@pytest.mark.parametrize("n,m", [(20, 40), (40, 80), (64, 128), (70, 140), (6, 5)])
def test_cov_basis_both_routes_with_excited_initial_bits(n, m):
    """K2 for a basis state with ones in it: the shared-memory kernel (narrow panels) and the DMMA GEMM route
    (P = E^T diag(s) O, Lambda = P - P^T; wide panels) against the oracle's R^T Lambda_0 R on a random column subset."""
    from matchcake_b200 import ops as mops

    rng = np.random.default_rng(n)
    bits = rng.integers(0, 2, n)
    r = gates.random_sptm(n, seed=n, batch_size=2)
    cols = rng.permutation(2 * n)[:m]
    got = mops.cov_basis(dev(r[:, :, cols]), n, torch.as_tensor(bits, dtype=torch.int8)).cpu().numpy()
    ref = cv.evolve(cv.covariance_matrix(cv.amplitudes_from_basis_state(bits)), r)[:, cols][:, :, cols]
    np.testing.assert_allclose(got, ref, rtol=RTOL, atol=1e-12)


@pytest.mark.parametrize("n", [2, 5, 16, 33, 64])
@pytest.mark.parametrize("m", [1, 3, 8])
def test_register_resident_chain_matches_general_kernel(n, m):
    """Nearest-neighbour circuits, narrow panels: the register-resident kernel (circuit_brick.cu) against the oracle and,
    bit for bit, against the shared-memory kernel (same arithmetic, different data movement)."""
    from matchcake_b200 import ops as mops

    rng = np.random.default_rng(100 * n + m)
    B = 3
    blk = gates.random_sptm(2, seed=n)
    blk_b = gates.random_sptm(2, seed=n + 1, batch_size=B)
    kinds = ["SptmCompRxRx", "SptmCompRyRy", "SptmCompRzRz"]
    ops = []
    for layer in range(7):
        for w in range(layer % 2, n - 1, 2):
            pick = (layer + w) % 6
            if pick < 3:
                ops.append(Op(kinds[pick], (w, w + 1), params=rng.uniform(0, 6, (B, 2))))
            elif pick == 3:
                ops.append(Op("SptmFSwap", (w, w + 1)))
            elif pick == 4:
                ops.append(Op("Sptm", (w, w + 1), matrix=blk))
            else:
                ops.append(Op("Sptm", (w, w + 1), matrix=blk_b))
    circ, params = compiled(ops, n, B)
    assert circ._struct.nearest_neighbour == 1
    cols = torch.as_tensor(rng.permutation(2 * n)[:m].astype(np.int32))
    ref = gates.chain(ops, n, contraction=None)[:, :, cols.numpy()]
    fast = mops.circuit_columns(circ, params, cols).cpu().numpy()
    np.testing.assert_allclose(fast, ref, rtol=RTOL, atol=ATOL)
    circ._struct.nearest_neighbour = 0
    general = mops.circuit_columns(circ, params, cols).cpu().numpy()
    np.testing.assert_array_equal(fast, general)


@pytest.mark.parametrize("n,n_features", [(4, 4), (6, 12), (32, 64), (128, 128)])
def test_kernel_ansatz_sptm(n, n_features):
    """FermionicPQCKernel ansatz (affine map folded into the kernel) vs the oracle's _x_to_sptm."""
    from matchcake_b200 import ops as mops

    rng = np.random.default_rng(n)
    B = 4 if n >= 32 else 9
    x = rng.random((B, n_features))
    k = kernels.FermionicPQCKernelOracle(n_qubits=n, rotations="X,Z").fit(x)
    ref = k.x_to_sptm(x)
    ops = k.ansatz(x)
    g, params, consts = ops_to_circuit(ops, n, B)
    circ = mops.CompiledCircuit(n, g, params.shape[1])
    out = mops.circuit_columns(circ, dev(params)).cpu().numpy()
    np.testing.assert_allclose(out, ref, rtol=RTOL, atol=ATOL)


# ------------------------------------------------------------------ K2 + read-out
@pytest.mark.parametrize("n,k", [(4, 4), (6, 3), (16, 16), (20, 5)])
def test_cov_basis_and_probs(n, k):
    from matchcake_b200 import ops as mops

    rng = np.random.default_rng(n + k)
    B = 6
    ops = readme_ops(n, rng.uniform(0, 2 * np.pi, (B, 2)), matchgate=False) + \
        [Op("SptmCompRxRx", (w, w + 1), params=rng.uniform(0, 6, (B, 2))) for w in range(1, n - 1, 2)]
    r = gates.chain(ops, n, contraction=None)
    bits = rng.integers(0, 2, n)
    wires = rng.permutation(n)[:k]
    maj = readout.majorana_indices_of_wires(wires)
    circ, params = compiled(ops, n, B)
    x = mops.circuit_columns(circ, params, torch.as_tensor(maj, dtype=torch.int32))
    cov = mops.cov_basis(x, n, torch.as_tensor(bits))
    cov_ref = cv.evolve(cv.covariance_matrix(cv.amplitudes_from_basis_state(bits)), r)
    np.testing.assert_allclose(cov.cpu().numpy(), cov_ref[:, maj[:, None], maj[None, :]], rtol=RTOL, atol=ATOL)
    if k <= 5:
        states = readout.states_to_binary(np.arange(2 ** k), k)
    else:
        states = rng.integers(0, 2, (11, k))
    p = mops.probs(cov, torch.as_tensor(np.ascontiguousarray(states))).cpu().numpy()
    # reference default path (complex lookup table) for basis-state inputs
    ref = readout.probs_lookup_table(r, bits, states, wires).T
    np.testing.assert_allclose(p, ref, rtol=RTOL, atol=ATOL)
    if k <= 5:
        pn = mops.probs(cov, torch.as_tensor(np.ascontiguousarray(states)), normalize=True).cpu().numpy()
        np.testing.assert_allclose(pn, readout.analytic_probability(r, cv.amplitudes_from_basis_state(bits), wires),
                                   rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("n,k,product", [(4, 4, False), (6, 3, True), (10, 8, False), (12, 10, True), (12, 11, False), (5, 1, True)])
def test_probs_all_tree(n, k, product):
    """All 2^k outcomes at once (shared elimination tree) == the reference's per-outcome formula, normalised or not."""
    from matchcake_b200 import ops as mops

    rng = np.random.default_rng(100 * n + k)
    B = 5
    r = gates.random_sptm(n, seed=n + k, batch_size=B)
    if product:
        amp = rng.standard_normal((n, 2)) + 1j * rng.standard_normal((n, 2))
        amp /= np.linalg.norm(amp, axis=1, keepdims=True)
    else:
        amp = cv.amplitudes_from_basis_state(rng.integers(0, 2, n))
    wires = rng.permutation(n)[:k]
    maj = readout.majorana_indices_of_wires(wires)
    cov_ref = cv.evolve(cv.covariance_matrix(amp), r)[:, maj[:, None], maj[None, :]]
    states = readout.states_to_binary(np.arange(2 ** k), k)
    ref = np.stack([readout.probs_product_state(cv.evolve(cv.covariance_matrix(amp), r)[b], states, wires) for b in range(B)])
    out = mops.probs_all(dev(cov_ref)).cpu().numpy()
    np.testing.assert_allclose(out, ref, rtol=RTOL, atol=ATOL)
    outn = mops.probs_all(dev(cov_ref), normalize=True).cpu().numpy()
    np.testing.assert_allclose(outn, ref / ref.sum(axis=1, keepdims=True), rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(out, mops.probs(dev(cov_ref), torch.as_tensor(states.copy())).cpu().numpy(),
                               rtol=RTOL, atol=ATOL)


def test_probs_all_parity_forbidden_outcomes_are_exact_zeros():
    """Basis-state input measured on ALL wires: half of the outcomes violate fermion parity -> exactly 0."""
    from matchcake_b200 import ops as mops

    n = 6
    rng = np.random.default_rng(7)
    ops = readme_ops(n, rng.uniform(0, 6, (3, 2)), matchgate=False)
    r = gates.chain(ops, n, contraction=None)
    bits = np.array([1, 0, 0, 1, 0, 0])
    cov_ref = cv.evolve(cv.covariance_matrix(cv.amplitudes_from_basis_state(bits)), r)
    out = mops.probs_all(dev(cov_ref), normalize=True).cpu().numpy()
    ref = readout.analytic_probability(r, cv.amplitudes_from_basis_state(bits), np.arange(n))
    np.testing.assert_allclose(out, ref, rtol=RTOL, atol=ATOL)
    odd = np.array([bin(q).count("1") % 2 for q in range(2 ** n)]) != bits.sum() % 2
    assert np.all(out[:, odd] == 0.0)


def test_fused_expval_projector_golden_and_batch():
    from matchcake_b200 import ops as mops

    g = G["joss_expval"]
    circ, params = compiled(readme_ops(4, np.array([g["params"]])), 4, 1)
    v = mops.circuit_expval_projector(circ, params).item()
    assert f"{v:.4f}" == g["printed"]
    rng = np.random.default_rng(1)
    for n, B in [(4, 33), (16, 64), (7, 5)]:
        p = rng.uniform(0, 2 * np.pi, (B, 2))
        ops = readme_ops(n, p)
        bits, tgt = rng.integers(0, 2, n), rng.integers(0, 2, n)
        if n == 16:
            bits = tgt = np.zeros(n, int)  # config C2
        r = gates.chain(ops, n)
        ref = readout.probs_lookup_table(r, bits, tgt, np.arange(n))
        circ, params = compiled(ops, n, B)
        out = mops.circuit_expval_projector(circ, params, torch.as_tensor(bits), torch.as_tensor(tgt)).cpu().numpy()
        np.testing.assert_allclose(out, ref, rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("route", [1, 2])
@pytest.mark.parametrize("n", [2, 5, 8, 13, 16, 20, 32])
def test_fused_expval_projector_routes_random_circuits(n, route):
    """Both algorithms behind mcb200_circuit_expval_projector (1: X = R columns + DMMA conjugation, 2: covariance
    evolved gate by gate) on random circuits: every rotation kind in runs of 1-4 on one wire pair, long-range fSWAPs in
    both orientations, shared and per-instance explicit 4x4 blocks, excited initial and target bits, ragged batch."""
    from matchcake_b200 import _lib
    from matchcake_b200 import ops as mops

    rng = np.random.default_rng(100 * n + route)
    B = 37
    names = ["SptmCompRxRx", "SptmCompRyRy", "SptmCompRzRz"]
    ops = []
    for _ in range(3 * n + 4):
        k = int(rng.integers(0, 6)) if n > 2 else int(rng.integers(0, 3))
        if k < 3:
            w = int(rng.integers(0, n - 1))
            for _ in range(int(rng.integers(1, 5))):
                ops.append(Op(names[int(rng.integers(0, 3))], (w, w + 1), params=rng.uniform(0, 2 * np.pi, (B, 2))))
        elif k < 5:
            w0 = int(rng.integers(0, n - 1))
            w1 = int(rng.integers(w0 + 1, n))
            ops.append(Op("SptmFSwap", (w0, w1) if rng.random() < 0.5 else (w1, w0)))
        else:
            w = int(rng.integers(0, n - 1))
            batched = rng.random() < 0.5
            blk = gates.random_sptm(2, seed=int(rng.integers(1 << 30)), batch_size=B if batched else None)
            ops.append(Op("Sptm", (w, w + 1), matrix=blk))
    bits, tgt = rng.integers(0, 2, n), rng.integers(0, 2, n)
    tgt[rng.integers(0, n)] ^= (bits.sum() + tgt.sum()) % 2  # same parity: non-trivial probability
    r = gates.chain(ops, n, contraction=None)
    ref = readout.probs_lookup_table(r, bits, tgt, np.arange(n))
    circ, params = compiled(ops, n, B)
    try:
        _lib.set_expval_route(route)
        out = mops.circuit_expval_projector(circ, params, torch.as_tensor(bits), torch.as_tensor(tgt)).cpu().numpy()
    finally:
        _lib.set_expval_route(0)
    np.testing.assert_allclose(out, ref, rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("route", [1, 2])
def test_fused_expval_projector_routes_shared_parameters(route):
    """Config C2's structure (every gate reads parameter columns 0, 1; affine map on the angles) on both routes, batch
    sizes around the group size of the covariance route's sincos pass."""
    from matchcake_b200 import _lib
    from matchcake_b200 import ops as mops

    n = 16
    rng = np.random.default_rng(5 + route)
    specs = []
    for w in range(0, n - 1, 2):
        for kind in (mops.GATE_RXRX, mops.GATE_RYRY, mops.GATE_RZRZ):
            specs.append(mops.GateSpec(kind, w, w + 1, 0))
    for w in range(1, n - 1, 2):
        specs.append(mops.GateSpec(mops.GATE_FSWAP, w, w + 1))
    scale, bias = np.array([np.pi, 0.5]), np.array([0.3, -0.2])
    for B in (1, 7, 8, 9, 33, 1000):
        p = rng.uniform(0, 2, (B, 2))
        circ = mops.CompiledCircuit(n, specs, 2, scale=scale, bias=bias)
        try:
            _lib.set_expval_route(route)
            out = mops.circuit_expval_projector(circ, dev(p)).cpu().numpy()
        finally:
            _lib.set_expval_route(0)
        zeros = np.zeros(n, int)
        ref = readout.probs_lookup_table(gates.chain(readme_ops(n, bias + scale * p), n), zeros, zeros, np.arange(n))
        np.testing.assert_allclose(out, ref, rtol=RTOL, atol=ATOL)


def test_dense_path_tutorial_golden():
    """Random dense SPTM, product-state inputs: probabilities and Pauli-sum expectation values of the tutorials."""
    from matchcake_b200 import ops as mops

    r = gates.random_sptm(3, seed=0)
    th = np.pi / 3
    mixed = np.array([[1.0, 1.0], [1.0, 0.0], [np.cos(th / 2), np.sin(th / 2)]], dtype=complex)
    mixed /= np.linalg.norm(mixed, axis=1, keepdims=True)
    plus = np.full((3, 2), 1 / np.sqrt(2), dtype=complex)
    states = readout.states_to_binary(np.arange(8), 3)
    h = G["hamiltonian"]
    for amp, pkey, ekey in [(cv.amplitudes_from_basis_state([1, 0, 1]), "probs_basis_101", "expval_basis_101"),
                            (plus, "probs_plus", "expval_plus"), (mixed, "probs_mixed", "expval_mixed")]:
        cov = mops.cov_dense(dev(r), dev(cv.covariance_matrix(amp)))
        p = mops.probs(cov, torch.as_tensor(np.ascontiguousarray(states)), normalize=True).cpu().numpy()[0]
        np.testing.assert_array_equal(np.round(p, 6) + 0.0, np.asarray(G[pkey]["value"]))
        ext0 = cv.extended_covariance_matrix(cv.covariance_matrix(amp), cv.displacement_vector(amp))
        ext = mops.cov_dense(dev(r), dev(ext0))
        np.testing.assert_allclose(ext.cpu().numpy()[0], cv.evolved_extended_covariance(amp, r), rtol=RTOL, atol=ATOL)
        total = None
        for sector, (idx, coeffs) in readout.pauli_sum_structure(h["pauli_words"], h["coeffs"], 3).items():
            feats = mops.sector_features(ext, torch.as_tensor(idx, dtype=torch.int32))
            total = mops.rowdot(feats, dev(coeffs), out=total)
        assert round(total.item(), 6) == h[ekey]


@pytest.mark.parametrize("n,terms", [(5, 40), (9, 25)])
def test_sector_features_random_pauli_sums(n, terms):
    from matchcake_b200 import ops as mops

    rng = np.random.default_rng(n)
    B = 4
    r = gates.random_sptm(n, seed=n, batch_size=B)
    amp = rng.standard_normal((n, 2)) + 1j * rng.standard_normal((n, 2))
    amp /= np.linalg.norm(amp, axis=1, keepdims=True)
    words = ["".join(rng.choice(list("IXYZ"), n)) for _ in range(terms)]
    coeffs = rng.standard_normal(terms)
    ext_ref = cv.evolved_extended_covariance(amp, r)
    ref = readout.expval_pauli_sum(ext_ref, words, coeffs)
    ext0 = cv.extended_covariance_matrix(cv.covariance_matrix(amp), cv.displacement_vector(amp))
    ext = mops.cov_dense(dev(r), dev(ext0))
    total = None
    for sector, (idx, sc) in readout.pauli_sum_structure(words, coeffs, n).items():
        feats = mops.sector_features(ext, torch.as_tensor(idx.reshape(len(sc), sector), dtype=torch.int32))
        total = mops.rowdot(feats, dev(sc), out=total)
    np.testing.assert_allclose(total.cpu().numpy(), ref, rtol=RTOL, atol=1e-12)


@pytest.mark.parametrize("m", [4, 8, 12, 16, 24, 32, 40, 48, 64, 66, 80, 100, 128, 130, 192, 256])
def test_pfaffian_static_path_hand_over(m):
    """The tile Pfaffian eliminates pairs (2s, 2s+1) in place while the natural pivot passes its guard and hands the
    remaining Schur complement to the pivoted routine otherwise: force the hand-over at every possible pair (zero and
    tiny natural pivots, first and second pair of a double step) and compare with the oracle's signed Pfaffian."""
    from matchcake_b200 import ops as mops

    rng = np.random.default_rng(m)
    mats = []
    for s in range(m // 2):
        for eps in (0.0, 1e-9):
            a = random_skew(rng, (m, m))
            a[2 * s, 2 * s + 1], a[2 * s + 1, 2 * s] = eps, -eps
            mats.append(a)
    # pivots that vanish only AFTER earlier eliminations (the natural pivot of the Schur complement is zero); not for the
    # last pair: that matrix is singular and its "Pfaffian" is rounding noise in every implementation
    for s in range(1, m // 2 - 1):
        a = random_skew(rng, (m, m))
        lead = a[:2 * s, :2 * s]
        b12 = a[:2 * s, 2 * s:2 * s + 2]
        corr = (b12.T @ np.linalg.inv(lead) @ b12)[0, 1]   # Schur complement entry = a[2s][2s+1] + corr
        a[2 * s, 2 * s + 1] = -corr
        a[2 * s + 1, 2 * s] = corr
        mats.append(a)
    a_np = np.stack(mats)
    ref = pfaffian.pfaffian_signed(a_np)
    out = mops.pfaffian(dev(a_np), signed=True).cpu().numpy()
    # these matrices are ill conditioned on purpose; two different pivot orders agree to ~1e-9 at m = 256
    np.testing.assert_allclose(out, ref, rtol=1e-9 if m <= 128 else 1e-8, atol=1e-12 * np.abs(ref).max())


@pytest.mark.parametrize("n,terms,B", [(5, 40, 4), (9, 200, 3), (4, 300, 70)])
def test_pauli_sum_small_fused_matches_oracle(n, terms, B):
    """Mixed-size sectors (0, 2, 4, 6) in ONE launch; longer Majorana strings accumulate through the sector path."""
    from matchcake_b200 import ops as mops

    rng = np.random.default_rng(100 + n)
    r = gates.random_sptm(n, seed=n, batch_size=B)
    amp = rng.standard_normal((n, 2)) + 1j * rng.standard_normal((n, 2))
    amp /= np.linalg.norm(amp, axis=1, keepdims=True)
    # mostly short strings (nearest-neighbour style) plus a few arbitrary ones
    words = []
    for t in range(terms):
        w = ["I"] * n
        if t % 5 == 4:
            w = list(rng.choice(list("IXYZ"), n))
        else:
            q = int(rng.integers(0, n - 1))
            w[q] = str(rng.choice(list("XYZ")))
            if t % 2:
                w[q + 1] = str(rng.choice(list("XYZ")))
        words.append("".join(w))
    words[0] = "I" * n
    coeffs = rng.standard_normal(terms)
    ref = readout.expval_pauli_sum(cv.evolved_extended_covariance(amp, r), words, coeffs)
    ext0 = cv.extended_covariance_matrix(cv.covariance_matrix(amp), cv.displacement_vector(amp))
    ext = mops.cov_dense(dev(r), dev(ext0))
    structure = readout.pauli_sum_structure(words, coeffs, n)
    sizes, flat, scal = [], [], []
    for sector, (idx, sc) in structure.items():
        if sector <= 6:
            sizes += [sector] * len(sc)
            flat.append(np.asarray(idx, dtype=np.int32).reshape(-1))
            scal.append(sc)
    assert {0, 2, 4} <= set(sizes)
    term_ptr = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    total = mops.pauli_sum_small(ext, torch.as_tensor(term_ptr), torch.as_tensor(np.concatenate(flat)),
                                 torch.as_tensor(np.concatenate(scal)))
    for sector, (idx, sc) in structure.items():
        if sector > 6:
            feats = mops.sector_features(ext, torch.as_tensor(idx.reshape(len(sc), sector), dtype=torch.int32))
            total = mops.rowdot(feats, dev(sc), out=total)
    np.testing.assert_allclose(total.cpu().numpy(), ref, rtol=RTOL, atol=1e-12)


@pytest.mark.parametrize("n,batch", [(6, 4), (8, 3), (32, 5), (64, 500), (70, 2), (128, 2)])
@pytest.mark.parametrize("ta,tb", [(False, False), (False, True), (True, False), (True, True)])
def test_sptm_matmul(n, batch, ta, tb):
    from matchcake_b200 import ops as mops

    rng = np.random.default_rng(n)
    a, b = rng.standard_normal((batch, n, n)), rng.standard_normal((batch, n, n))
    ref = (np.swapaxes(a, 1, 2) if ta else a) @ (np.swapaxes(b, 1, 2) if tb else b)
    out = mops.sptm_matmul(dev(a), dev(b), ta, tb).cpu().numpy()
    np.testing.assert_allclose(out, ref, rtol=1e-12, atol=1e-12)
    shared = mops.sptm_matmul(dev(a), dev(b[0]), ta, tb).cpu().numpy()
    np.testing.assert_allclose(shared, (np.swapaxes(a, 1, 2) if ta else a) @ (b[0].T if tb else b[0]), rtol=1e-12, atol=1e-12)


# ------------------------------------------------------------------ Gram
@pytest.mark.parametrize("n,shape", [(4, (13, 13)), (4, (9, 14)), (4, (14, 9)), (16, (10, 10)), (32, (6, 6)), (40, (5, 6)), (64, (4, 4)), (70, (3, 4)), (128, (3, 3))])
def test_gram_matches_reference_formulation(n, shape):
    from matchcake_b200 import ops as mops

    rng = np.random.default_rng(n)
    nf = 2 * n if n <= 32 else n
    x0, x1 = rng.random((shape[0], nf)), rng.random((shape[1], nf))
    k = kernels.FermionicPQCKernelOracle(n_qubits=n, rotations="X,Z").fit(x0)
    ref = k.gram(x0, x1, float32_store=False)
    lam0 = cv.covariance_matrix(cv.amplitudes_from_basis_state(np.zeros(n, int)))
    c0, c1 = cv.evolve(lam0, k.x_to_sptm(x0)), cv.evolve(lam0, k.x_to_sptm(x1))
    out = mops.gram(dev(c0), dev(c1), scale=2.0 ** -n).cpu().numpy()
    np.testing.assert_allclose(out, ref, rtol=RTOL, atol=ATOL)
    full = mops.gram(dev(c0), dev(c1), scale=2.0 ** -n, mode=mops.GRAM_FULL).cpu().numpy()
    idx = np.stack(np.meshgrid(np.arange(shape[0]), np.arange(shape[1]), indexing="ij"), -1).reshape(-1, 2)
    np.testing.assert_allclose(full.reshape(-1), k.pair_expvals(k.x_to_sptm(x0), k.x_to_sptm(x1), idx), rtol=RTOL, atol=ATOL)
    # row-sharded triangle + symmetrize == single call
    tri = torch.zeros(shape, dtype=torch.float64, device="cuda")
    h = shape[0] // 2
    mops.gram(dev(c0), dev(c1), scale=2.0 ** -n, mode=mops.GRAM_TRIANGLE, row_range=(0, h), out=tri)
    mops.gram(dev(c0), dev(c1), scale=2.0 ** -n, mode=mops.GRAM_TRIANGLE, row_range=(h, shape[0]), out=tri)
    np.testing.assert_array_equal(mops.gram_symmetrize(tri).cpu().numpy(), out)
    # the same with rows-only buffers (views into the final matrix; nothing zero-filled) -- what the multi-GPU build does
    final = torch.full(shape, float("nan"), dtype=torch.float64, device="cuda")
    mops.gram(dev(c0), dev(c1), scale=2.0 ** -n, mode=mops.GRAM_TRIANGLE, row_range=(0, h), out=final[:h], rows_only=True)
    mops.gram(dev(c0), dev(c1), scale=2.0 ** -n, mode=mops.GRAM_TRIANGLE, row_range=(h, shape[0]), out=final[h:], rows_only=True)
    np.testing.assert_array_equal(mops.gram_symmetrize(final).cpu().numpy(), out)
    with pytest.raises(ValueError):
        mops.gram(dev(c0), dev(c1), scale=1.0, row_range=(0, h), out=final[:h], rows_only=True)  # REFERENCE mirrors outside


def test_gram_full_size_cells_whose_static_guard_fails():
    """64 x 64 cells go through the branch-free static elimination (pfaffian_flat.cuh): a failed block-pivot guard is
    only recorded and the cell is redone from its inputs with partial pivoting.  Physical inputs (covariances
    Q J Q^T of pure Gaussian states of one parity -- the guard's growth bound assumes them): Lambda_j = S Lambda_i S^T
    with S = P_4 + R', P_4 X P_4^T = -X for the leading block X of Lambda_i, makes the leading 4 x 4 block of
    Lambda_i + Lambda_j vanish while the sum stays well conditioned; a further rotation by exp(eps K) makes the block
    tiny instead (eps = 1e-4: small enough to be a bad pivot, large enough to pass the guard).  Ordinary cells sit next
    to the forced ones.  Every entry against the oracle's Pfaffian of the sum."""
    from scipy.linalg import expm, schur

    from matchcake_b200 import ops as mops

    d, n = 64, 12
    rng = np.random.default_rng(64)
    j0 = np.kron(np.eye(d // 2), np.array([[0.0, 1.0], [-1.0, 0.0]]))

    def rand_so(k):
        q, _ = np.linalg.qr(rng.standard_normal((k, k)))
        if np.linalg.det(q) < 0:
            q[:, 0] = -q[:, 0]
        return q

    def flip4(x):  # real Schur form x = O (x1 J + x2 J) O^T  ->  P = O diag(1, -1, 1, -1) O^T, det P = +1
        _, o = schur(x, output="real")
        return o @ np.diag([1.0, -1.0, 1.0, -1.0]) @ o.T

    a = np.stack([q @ j0 @ q.T for q in (rand_so(d) for _ in range(n))])
    b = np.stack([q @ j0 @ q.T for q in (rand_so(d) for _ in range(n))])
    forced = {}
    for t, eps in enumerate((0.0, 0.0, 1e-12, 1e-9, 1e-6, 1e-4)):
        i, j = 2 * t, 2 * t + 1
        r = np.eye(d)
        r[:4, :4] = flip4(a[i][:4, :4])
        r[4:, 4:] = rand_so(d - 4)
        r = r @ expm(eps * random_skew(rng, (d, d)))
        b[j] = r @ a[i] @ r.T
        forced[(i, j)] = eps
    a, b = (a - a.transpose(0, 2, 1)) / 2, (b - b.transpose(0, 2, 1)) / 2
    out = mops.gram(dev(a), dev(b), scale=1.0, mode=mops.GRAM_FULL).cpu().numpy()
    ref = np.abs(pfaffian.pfaffian_signed((a[:, None] + b[None, :]).reshape(-1, d, d))).reshape(n, n)
    assert ref.min() > 1e-3                                            # nothing here is (near) singular
    loose = np.zeros((n, n), bool)
    for (i, j), eps in forced.items():
        assert np.abs((a[i] + b[j])[:4, :4]).max() <= max(8 * eps, 1e-14)
        # a small block may pass the guard with a growth of up to kStaticBlockGrowthPhysical = 1e5 (times
        # the unit round-off: 1e-11 .. 1e-10); real cells reach 6e4 at the 99.9 % quantile (pfaffian_tiles.cuh)
        loose[i, j] = eps > 0.0
    np.testing.assert_allclose(out[~loose], ref[~loose], rtol=RTOL, atol=0)
    np.testing.assert_allclose(out[loose], ref[loose], rtol=1e-9, atol=0)


def _haar_covariances(rng, n, d):
    j0 = np.kron(np.eye(d // 2), np.array([[0.0, 1.0], [-1.0, 0.0]]))
    out = []
    for _ in range(n):
        q, _ = np.linalg.qr(rng.standard_normal((d, d)))
        if np.linalg.det(q) < 0:
            q[:, 0] = -q[:, 0]
        lam = q @ j0 @ q.T
        out.append((lam - lam.T) / 2)
    return np.stack(out)


@pytest.mark.parametrize("d", [32, 64])
def test_gram_of_generic_gaussian_states_pivot_bound(d):
    """Covariances of Haar-random pure Gaussian states (no light-cone structure): the growth bound of the static block
    pivots that is exact on the reference's kernel ansatz (1e5) lets errors of 1e-9 .. 1e-8 through in a few cells per
    thousand; the strict bound (1e3) keeps every cell within 1e-10 of the oracle; "auto" (the default) must detect the
    difference on its sample of the cells and take the strict route.  On ansatz covariances "auto" keeps the fast one."""
    from matchcake_b200 import ops as mops

    n = 260  # 67 600 cells: above the size below which "auto" is strict anyway
    rng = np.random.default_rng(d)
    cov = _haar_covariances(rng, n, d)
    c = dev(cov)
    strict = mops.gram(c, c, 1.0, mode=mops.GRAM_FULL, pivoting="strict").cpu().numpy()
    fast = mops.gram(c, c, 1.0, mode=mops.GRAM_FULL, pivoting="fast").cpu().numpy()
    # oracle on the 300 cells where the two routes differ most and on 3000 random ones (the full set takes a minute of LU)
    diff = np.abs(fast - strict) / np.abs(strict)
    np.fill_diagonal(diff, 0.0)
    sel = np.concatenate([np.argsort(diff.reshape(-1))[-300:], rng.choice(n * n, 3000, replace=False)])
    sel = sel[sel // n != sel % n]
    ii, jj = sel // n, sel % n
    ref = np.abs(pfaffian.pfaffian_signed(cov[ii] + cov[jj]))
    np.testing.assert_allclose(strict[ii, jj], ref, rtol=RTOL, atol=1e-13 * ref.max())
    np.testing.assert_allclose(fast[ii, jj], ref, rtol=1e-6, atol=1e-13 * ref.max())   # what "fast" means on such data
    assert mops._gram_growth_for(c, c, n * n, "auto") == mops.GRAM_GROWTH_STRICT
    auto = mops.gram(c, c, 1.0, mode=mops.GRAM_FULL).cpu().numpy()
    np.testing.assert_array_equal(auto, strict)
    assert float(mops._lib.load().mcb200_get_gram_pivot_growth()) == mops.GRAM_GROWTH_FAST  # option restored
    # the kernel ansatz: fast and strict agree, "auto" keeps the fast bound
    nq = d // 2
    x = rng.random((n, 2 * nq))
    k = kernels.FermionicPQCKernelOracle(n_qubits=nq, rotations="X,Z").fit(x)
    lam0 = cv.covariance_matrix(cv.amplitudes_from_basis_state(np.zeros(nq, int)))
    ca = dev(cv.evolve(lam0, k.x_to_sptm(x)))
    assert mops._gram_growth_for(ca, ca, n * n, "auto") == mops.GRAM_GROWTH_FAST
    with pytest.raises(ValueError):
        mops.gram(ca, ca, 1.0, pivoting="sometimes")


@pytest.mark.parametrize("shape", [(1, 1), (33, 33), (70, 45), (45, 70), (200, 97), (64, 64)])
def test_gram_symmetrize_tiles(shape):
    """GramMatrix.symmetrize_ / _fill_diagonal_ (gram_matrix.py:127-161) by the tiled transpose kernel."""
    from matchcake_b200 import ops as mops

    n0, n1 = shape
    a = np.random.default_rng(n0 * 1000 + n1).random(shape)
    q = min(n0, n1)
    ref = a.copy()
    sq = ref[:q, :q]
    src = np.tril(sq, -1) if n0 >= n1 else np.triu(sq, 1)               # evaluated triangle (gram_matrix.py:191-195)
    sq[...] = src + src.T + np.eye(q)
    # a non-contiguous view (row block of a wider matrix) exercises the leading dimension
    wide = torch.zeros((n0, n1 + 5), dtype=torch.float64, device="cuda")
    wide[:, :n1] = dev(a)
    got = mops.gram_symmetrize(wide[:, :n1])
    np.testing.assert_array_equal(got.cpu().numpy(), ref)
    assert float(wide[:, n1:].abs().max()) == 0.0


def test_gram_identical_rows_is_all_ones():
    """tests/test_ml/test_kernels/test_fermionic_pqc_kernel.py:111-115."""
    from matchcake_b200 import ops as mops

    n = 8
    x = np.full((5, 2 * n), 0.37)
    k = kernels.FermionicPQCKernelOracle(n_qubits=n, rotations="Y,Z").fit(x)
    lam0 = cv.covariance_matrix(cv.amplitudes_from_basis_state(np.zeros(n, int)))
    c = dev(cv.evolve(lam0, k.x_to_sptm(x)))
    np.testing.assert_allclose(mops.gram(c, c, scale=2.0 ** -n).cpu().numpy(), 1.0, rtol=1e-12)


def test_errors_are_loud():
    from matchcake_b200 import ops as mops
    from matchcake_b200._lib import Mcb200Error

    with pytest.raises(Mcb200Error):
        mops.pfaffian(torch.zeros((2, 4, 4), dtype=torch.float64))  # CPU tensor: no CPU path
    with pytest.raises(ValueError):
        mops.CompiledCircuit(3, [mops.GateSpec(mops.GATE_RXRX, 2, 3, 0)], 2)
    with pytest.raises(Mcb200Error):  # Pauli sectors beyond 128 Majorana operators are not supported
        mops.sector_features(torch.zeros((1, 140, 140), dtype=torch.float64, device="cuda"),
                             torch.arange(130, dtype=torch.int32, device="cuda")[None])
    with pytest.raises(Mcb200Error):  # the C entry point itself stops at 64 measured wires (ops.probs composes beyond)
        from matchcake_b200 import _lib
        cov = torch.zeros((1, 140, 140), dtype=torch.float64, device="cuda")
        oc = torch.zeros((1, 70), dtype=torch.int8, device="cuda")
        out = torch.zeros((1, 1), dtype=torch.float64, device="cuda")
        _lib.check(_lib.load().mcb200_probs(cov.data_ptr(), oc.data_ptr(), 1, 1, 70, 0, out.data_ptr(), None))


def test_matchgate_sptm_kernel():
    """Generic matchgate -> SPTM on the GPU == trace formula of the reference (utils/__init__.py:561-596)."""
    from matchcake_b200 import ops as mops

    rng = np.random.default_rng(2)
    us = []
    for kind in ("XX", "YY", "ZZ"):
        us.append(gates.matchgate_comp_rotation(rng.uniform(0, 6, (6, 2)), kind))
    u = np.concatenate(us + [gates.matchgate_fswap()[None]], axis=0)
    u = np.einsum("bij,bjk->bik", u, u[::-1])  # products of matchgates on the same wires are matchgates
    ref = gates.enforce_real(gates.sptm_from_gate(u))
    out = mops.matchgate_sptm(torch.as_tensor(u).cuda()).cpu().numpy()
    np.testing.assert_allclose(out, ref, rtol=RTOL, atol=ATOL)
    one = mops.matchgate_sptm(torch.as_tensor(u[3]).cuda()).cpu().numpy()
    np.testing.assert_allclose(one, ref[3], rtol=RTOL, atol=ATOL)
    # Tr(U c_mu U^dag c_nu) is a trace of two Hermitian matrices, hence real for ANY U: the REAL_PART_ATOL guard of
    # the reference can only fire on NaNs here, and the kernel reports max |Im| = rounding noise
    generic = np.linalg.qr(rng.standard_normal((4, 4)) + 1j * rng.standard_normal((4, 4)))[0]
    assert np.isfinite(mops.matchgate_sptm(torch.as_tensor(generic).cuda()).cpu().numpy()).all()


@pytest.mark.parametrize("n,batch", [(1, None), (3, None), (7, 4), (40, 3)])
def test_product_state_covariance_kernel(n, batch):
    from matchcake_b200 import ops as mops

    rng = np.random.default_rng(n)
    shape = (n, 2) if batch is None else (batch, n, 2)
    amp = rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
    amp /= np.linalg.norm(amp, axis=-1, keepdims=True)
    cov = mops.product_state_cov(torch.as_tensor(amp).cuda()).cpu().numpy()
    np.testing.assert_allclose(cov, cv.covariance_matrix(amp), rtol=RTOL, atol=ATOL)
    ext = mops.product_state_cov(torch.as_tensor(amp).cuda(), extended=True).cpu().numpy()
    ref = cv.extended_covariance_matrix(cv.covariance_matrix(amp), cv.displacement_vector(amp))
    np.testing.assert_allclose(ext, ref, rtol=RTOL, atol=ATOL)


def test_pair_blocks_match_dense_covariance():
    """circuit_pair_blocks == the 4x4 diagonal blocks of the dense covariance, with a non-trivial initial state."""
    from matchcake_b200 import ops as mops

    rng = np.random.default_rng(9)
    n, B = 8, 6
    ops = []
    for w in range(0, n, 2):
        ops += [Op("SptmCompRyRy", (w, w + 1), params=rng.uniform(0, 6, (B, 2))), Op("SptmFSwap", (w, w + 1)),
                Op("SptmCompRzRz", (w, w + 1), params=rng.uniform(0, 6, (B, 2)))]
    bits = rng.integers(0, 2, n)
    r = gates.chain(ops, n, contraction=None)
    cov = cv.evolve(cv.covariance_matrix(cv.amplitudes_from_basis_state(bits)), r)
    circ, params = compiled(ops, n, B)
    blocks = mops.circuit_pair_blocks(circ, params, torch.arange(0, n, 2), torch.as_tensor(bits)).cpu().numpy()
    iu = np.triu_indices(4, 1)
    for c in range(n // 2):
        np.testing.assert_allclose(blocks[:, c], cov[:, 4 * c:4 * c + 4, 4 * c:4 * c + 4][:, iu[0], iu[1]], rtol=RTOL, atol=ATOL)
    off = cov.copy()
    for c in range(n // 2):
        off[:, 4 * c:4 * c + 4, 4 * c:4 * c + 4] = 0
    assert np.abs(off).max() < 1e-14  # the covariance really is block diagonal


@pytest.mark.parametrize("n,depth,cols", [(12, 5, [0, 1, 6, 7]), (16, 20, [30, 31]), (10, 6, [4, 5, 18, 19]), (64, 64, list(range(16)))])
def test_light_cone_pruning_keeps_the_columns(n, depth, cols):
    """circuit_columns(prune=True) launches only the gates inside the backward light cone of the panel: same columns as
    the full gate list (zeros may differ in sign), fewer gates; long-range fSWAPs and explicit blocks included."""
    from matchcake_b200 import ops as mops

    rng = np.random.default_rng(n + depth)
    kinds = (mops.GATE_RXRX, mops.GATE_RYRY, mops.GATE_RZRZ)
    g, off, consts = [], 0, []
    for l in range(depth):
        for q in range(l % 2, n - 1, 2):
            g.append(mops.GateSpec(kinds[(l + q) % 3], q, q + 1, off))
            off += 2
        if l % 4 == 1 and n <= 16:
            a, b = sorted(rng.choice(n, 2, replace=False))
            g.append(mops.GateSpec(mops.GATE_FSWAP, int(a), int(b), 0, mops.FLAG_TRANSPOSE if l % 8 == 1 else 0))
        if l % 5 == 2 and n <= 16:
            q = int(rng.integers(0, n - 1))
            g.append(mops.GateSpec(mops.GATE_BLOCK4_CONST, q, q + 1, len(consts)))
            consts.extend(np.linalg.qr(rng.standard_normal((4, 4)))[0].reshape(-1).tolist())
    circ = mops.CompiledCircuit(n, g, off, consts=np.asarray(consts) if consts else None)
    params = dev(rng.uniform(0, 2 * np.pi, (5, off)))
    sub, _ = circ.light_cone(cols)
    assert len(sub.gates) <= len(circ.gates)
    if depth < n // 2 or n == 64:
        assert len(sub.gates) < len(circ.gates)   # shallow against the register / panel at the edge: gates really drop out
    full = mops.circuit_columns(circ, params, cols, prune=False).cpu().numpy()
    pruned = mops.circuit_columns(circ, params, np.asarray(cols)).cpu().numpy()
    pruned_t = mops.circuit_columns(circ, params, torch.as_tensor(cols, device="cuda")).cpu().numpy()
    # pruning drops exact zeros only; panels of 9..32 columns take the column-major kernel, whose fused multiply-adds may
    # round differently from the row-major kernel's (1 ulp)
    np.testing.assert_allclose(pruned, full, rtol=1e-12, atol=1e-15)
    np.testing.assert_array_equal(pruned_t, pruned)
    np.testing.assert_allclose(full, mops.circuit_columns(circ, params)[:, :, cols].cpu().numpy(), rtol=1e-12, atol=1e-15)
    # gradients: the dropped gates get exact zeros
    p = params.clone().requires_grad_(True)
    w = dev(rng.standard_normal(full.shape))
    (mops.circuit_columns(circ, p, cols) * w).sum().backward()
    gp = p.grad.clone()
    p.grad = None
    (mops.circuit_columns(circ, p, cols, prune=False) * w).sum().backward()
    np.testing.assert_allclose(gp.cpu().numpy(), p.grad.cpu().numpy(), rtol=1e-12, atol=1e-13)


@pytest.mark.parametrize("n,depth,cols,B", [(12, 9, [2, 3, 6, 7], 7), (16, 20, list(range(10, 22)), 5), (64, 64, list(range(16)), 9),
                                            (40, 12, list(range(0, 32)), 3), (7, 5, [0, 1, 12, 13, 5], 4)])
def test_fused_circuit_cov_basis_matches_two_kernels_and_oracle(n, depth, cols, B):
    """mcb200_circuit_cov_basis (column-major panel kernel, covariance in the same block) == circuit_columns + cov_basis
    (row-major kernels) == oracle; excited initial bits, adjacent fSWAPs and explicit blocks in the stream, odd panel
    widths, the column kernel on its own (9..32 columns), and the unfused fall-back (long-range fSWAP)."""
    from matchcake_b200 import ops as mops

    rng = np.random.default_rng(n * depth)
    kinds = (mops.GATE_RXRX, mops.GATE_RYRY, mops.GATE_RZRZ)
    names = ("SptmCompRxRx", "SptmCompRyRy", "SptmCompRzRz")
    g, oops, off, consts = [], [], 0, []
    cols_p = []
    for l in range(depth):
        for q in range(l % 2, n - 1, 2):
            k = (l + q) % 3
            g.append(mops.GateSpec(kinds[k], q, q + 1, off))
            cols_p.append((names[k], q, off))
            off += 2
        if l % 3 == 1:
            q = int(rng.integers(0, n - 1))
            g.append(mops.GateSpec(mops.GATE_FSWAP, q, q + 1, 0, mops.FLAG_TRANSPOSE if l % 2 else 0))
            cols_p.append(("fswap", q, (q + 1, q) if l % 2 else (q, q + 1)))
        if l % 4 == 2:
            q = int(rng.integers(0, n - 1))
            blk = np.linalg.qr(rng.standard_normal((4, 4)))[0]
            g.append(mops.GateSpec(mops.GATE_BLOCK4_CONST, q, q + 1, len(consts)))
            consts.extend(blk.reshape(-1).tolist())
            cols_p.append(("block", q, blk))
    params = rng.uniform(0, 2 * np.pi, (B, off))
    for nm, q, o in cols_p:
        if nm == "fswap":
            oops.append(Op("SptmFSwap", o))
        elif nm == "block":
            oops.append(Op("Sptm", (q, q + 1), matrix=o))
        else:
            oops.append(Op(nm, (q, q + 1), params=params[:, o:o + 2]))
    circ = mops.CompiledCircuit(n, g, off, consts=np.asarray(consts))
    bits = rng.integers(0, 2, n)
    p = dev(params)
    fused = mops.circuit_cov_basis(circ, p, cols, torch.as_tensor(bits))
    x = mops.circuit_columns(circ, p, cols)
    two = mops.cov_basis(x, n, torch.as_tensor(bits))
    np.testing.assert_allclose(fused.cpu().numpy(), two.cpu().numpy(), rtol=1e-13, atol=1e-14)
    assert float((fused + fused.transpose(1, 2)).abs().max()) == 0.0
    # the column-major kernel against the row-major one (same arithmetic, fused multiply-adds may round differently)
    np.testing.assert_allclose(x.cpu().numpy(), mops.circuit_columns(circ, p)[:, :, cols].cpu().numpy(), rtol=1e-12, atol=1e-15)
    r = gates.chain(oops, n)
    amp = cv.amplitudes_from_basis_state(bits)
    ref = cv.evolve(cv.covariance_matrix(amp), r)[:, cols][:, :, cols]
    np.testing.assert_allclose(fused.cpu().numpy(), ref, rtol=RTOL, atol=ATOL)
    # a long-range fSWAP switches the fused entry point to its two-kernel path (workspace)
    if n <= 16:
        g2 = g + [mops.GateSpec(mops.GATE_FSWAP, 0, n - 1, 0)]
        circ2 = mops.CompiledCircuit(n, g2, off, consts=np.asarray(consts))
        f2 = mops.circuit_cov_basis(circ2, p, cols, torch.as_tensor(bits))
        t2 = mops.cov_basis(mops.circuit_columns(circ2, p, cols), n, torch.as_tensor(bits))
        np.testing.assert_allclose(f2.cpu().numpy(), t2.cpu().numpy(), rtol=1e-13, atol=1e-15)


def test_circuit_validate_accepts_good_and_rejects_malformed_descriptors():
    """mcb200_circuit_validate: what a non-Python host calls once per circuit structure (the kernels trust
    layer_ptr / wires; a malformed descriptor must be caught at the boundary, not corrupt shared memory)."""
    import ctypes as C

    from matchcake_b200 import _lib
    from matchcake_b200 import ops as mops
    from matchcake_b200._lib import Mcb200Error

    g = [mops.GateSpec(mops.GATE_RXRX, 0, 1, 0), mops.GateSpec(mops.GATE_RYRY, 2, 3, 2), mops.GateSpec(mops.GATE_FSWAP, 1, 2),
         mops.GateSpec(mops.GATE_FSWAP, 0, 3, 0, mops.FLAG_TRANSPOSE)]
    circ = mops.CompiledCircuit(4, g, 4)
    circ.validate()
    lib = _lib.load()

    def broken(mutate, match):
        c2 = mops.CompiledCircuit(4, g, 4)
        mutate(c2)
        with pytest.raises(Mcb200Error, match=match):
            c2.validate()

    def share_wire(c2):   # two gates on wire 1 in one layer
        c2.layer_ptr_dev.copy_(torch.tensor([0, 3, 3, 4], dtype=torch.int32))
    broken(share_wire, "share wire")

    def wire_out_of_range(c2):
        arr = c2.gates_dev.clone()
        arr[1, 2] = 7
        c2.gates_dev.copy_(arr)
    broken(wire_out_of_range, "outside")

    def param_overrun(c2):
        arr = c2.gates_dev.clone()
        arr[1, 3] = 3
        c2.gates_dev.copy_(arr)
    broken(param_overrun, "parameter columns")

    def bad_ptr(c2):
        c2.layer_ptr_dev.copy_(torch.tensor([0, 2, 1, 4], dtype=torch.int32))
    broken(bad_ptr, "monotone|share wire")

    def lying_flag(c2):
        c2._struct.nearest_neighbour = 1
    broken(lying_flag, "nearest_neighbour")
    assert lib.mcb200_circuit_validate(None) != 0
