"""Pins the CPU oracle against every known answer the reference publishes for the hot path
(tests/golden/reference_golden.json) and against the KATs of the reference's own test-suite."""
import json
import os

import numpy as np
import pytest

from oracle import covariance as cv
from oracle import gates, kernels, lookup_table, pfaffian, readout
from oracle.gates import Op

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_golden.json")))


def readme_circuit_ops(n, params, matchgate=True):
    """README.md:48-66."""
    names = ["CompRxRx", "CompRyRy", "CompRzRz"] if matchgate else ["SptmCompRxRx", "SptmCompRyRy", "SptmCompRzRz"]
    ops = []
    for w in range(0, n - 1, 2):
        for k in names:
            ops.append(Op(k, (w, w + 1), params=np.asarray(params)))
    for w in range(1, n - 1, 2):
        ops.append(Op("fSWAP" if matchgate else "SptmFSwap", (w, w + 1)))
    return ops


def mixed_amplitudes():
    th = np.pi / 3
    a = np.array([[1.0, 1.0], [1.0, 0.0], [np.cos(th / 2), np.sin(th / 2)]], dtype=complex)
    return a / np.linalg.norm(a, axis=1, keepdims=True)


PLUS = np.full((3, 2), 1.0 / np.sqrt(2), dtype=complex)


def test_joss_expval():
    g = G["joss_expval"]
    n = g["n_wires"]
    r = gates.chain(readme_circuit_ops(n, g["params"]), n)
    amp = cv.amplitudes_from_basis_state(np.zeros(n, int))
    p = readout.states_probability(r, amp, np.zeros(n, int), np.arange(n))
    assert f"{p:.4f}" == g["printed"]
    # real-covariance formulation == complex lookup-table formulation
    p2 = readout.probs_product_state(cv.evolve(cv.covariance_matrix(amp), r), np.zeros(n, int), np.arange(n))
    assert abs(p - p2) < 1e-13
    # SPTM-flavoured ops (closed forms) and no contraction give the same SPTM
    r2 = gates.chain(readme_circuit_ops(n, g["params"], matchgate=False), n, contraction=None)
    np.testing.assert_allclose(r, r2, atol=1e-14)


@pytest.mark.parametrize("key,amp", [("probs_basis_101", None), ("probs_plus", PLUS), ("probs_mixed", mixed_amplitudes())])
def test_probs_tutorial(key, amp):
    g = G[key]
    r = gates.random_sptm(g["n_wires"], seed=g["seed"])
    if amp is None:
        amp = cv.amplitudes_from_basis_state(g["input_bits"])
    p = readout.analytic_probability(r, amp, g["wires"])
    np.testing.assert_array_equal(np.round(p, g["decimals"]) + 0.0, np.asarray(g["value"]))
    assert abs(p.sum() - 1) < 1e-12


def test_probs_marginal_and_projector():
    g = G["probs_marginal_q1"]
    r = gates.random_sptm(3, seed=0)
    amp = cv.amplitudes_from_basis_state(g["input_bits"])
    np.testing.assert_array_equal(np.round(readout.analytic_probability(r, amp, g["wires"]), 6), g["value"])
    g = G["projector_101"]
    p = readout.states_probability(r, amp, g["target"], np.arange(3))
    assert round(float(p), 6) == g["value"]


def test_hamiltonian_expvals():
    g = G["hamiltonian"]
    r = gates.random_sptm(3, seed=0)
    for key, amp in [("expval_basis_101", cv.amplitudes_from_basis_state([1, 0, 1])), ("expval_plus", PLUS),
                     ("expval_mixed", mixed_amplitudes())]:
        e = readout.expval_pauli_sum(cv.evolved_extended_covariance(amp, r), g["pauli_words"], g["coeffs"])
        assert round(float(e), 6) == g[key]


def test_pfaffian_kats():
    k = G["pfaffian_kats"]
    assert pfaffian.signed_pfaffian(np.array(k["two_by_two"]["matrix"])) == pytest.approx(3.0, abs=1e-14)
    a, b, c, d, e, f = k["four_by_four"]["abcdef"]
    m = np.array([[0, a, b, c], [-a, 0, d, e], [-b, -d, 0, f], [-c, -e, -f, 0]])
    assert pfaffian.signed_pfaffian(m) == pytest.approx(a * f - b * e + c * d, abs=1e-13)
    assert pfaffian.sector_pfaffian_features(m, np.array([[0, 1, 2, 3]]))[0] == pytest.approx(8.0, abs=1e-13)
    pm = np.array(k["pivot_swap"]["matrix"])
    assert pfaffian.signed_pfaffian(pm) ** 2 == pytest.approx(np.linalg.det(pm), abs=1e-12)
    s = k["sector_2x2"]
    np.testing.assert_allclose(pfaffian.sector_pfaffian_features(np.array(s["cov"]), np.array(s["index_sets"])), s["value"])
    assert pfaffian.signed_pfaffian(np.zeros((0, 0))) == 1.0
    assert pfaffian.signed_pfaffian(np.zeros((3, 3))) == 0.0
    assert pfaffian.pfaffian(np.zeros((4, 4)), sign=True) == 0.0
    assert pfaffian.pfaffian(np.zeros((4, 4)), sign=False) ** 2 == pytest.approx(0.0, abs=1e-30)


@pytest.mark.parametrize("n,batch", [(2, None), (4, None), (2, 3), (4, 3), (8, 5), (16, 4), (64, 2)])
def test_pfaffian_squares_to_det(n, batch):
    rng = np.random.default_rng(n)
    m = rng.random((n, n) if batch is None else (batch, n, n))
    m = m - np.swapaxes(m, -1, -2)
    det = np.linalg.det(m)
    np.testing.assert_allclose(pfaffian.pfaffian(m, sign=True) ** 2, det, rtol=1e-9)
    np.testing.assert_allclose(pfaffian.pfaffian(m, sign=False) ** 2, np.abs(det), rtol=1e-9)


def test_jordan_wigner_table():
    for pauli, wires, idx, (re, im) in G["jordan_wigner_table"]["cases"]:
        mu, ph = readout.pauli_to_majorana(pauli, wires)
        np.testing.assert_array_equal(mu, idx)
        assert abs(ph - complex(re, im)) < 1e-15


def test_jordan_wigner_brute_force_n2():
    """P == phase * prod c_mu for all 16 two-qubit Pauli words (test_m_pfaffian_expval_strategy.py:121)."""
    import itertools

    paulis = {"I": gates.PAULI_I, "X": gates.PAULI_X, "Y": gates.PAULI_Y, "Z": gates.PAULI_Z}
    for chars in itertools.product("IXYZ", repeat=2):
        mu, ph = readout.pauli_to_majorana("".join(chars), [0, 1])
        mat = np.eye(4, dtype=complex)
        for i in mu:
            mat = mat @ gates.majorana(int(i), 2)
        np.testing.assert_allclose(ph * mat, np.kron(paulis[chars[0]], paulis[chars[1]]), atol=1e-14)


def test_lookup_table_kats():
    for case in G["lookup_table_kats"]["cases"]:
        t = 0.5 * np.array([[complex(*z) for z in row] for row in case["T_times_2"]])
        state = np.array([int(c) for c in case["state"]])
        obs = lookup_table.observable_matrix(t, state, np.array([[1]]), np.array([[0]]))
        np.testing.assert_allclose(obs.squeeze(), np.array(case["observable"]), atol=1e-12)


@pytest.mark.parametrize("n", [2, 3, 4])
def test_lookup_items(n):
    """tests/test_base/test_lookup_table.py:60-...: items against explicit products."""
    rng = np.random.default_rng(n)
    t = rng.random((n, 2 * n)) + 1j * rng.random((n, 2 * n))
    b = lookup_table.block_diagonal_matrix(n)
    it = lookup_table.lookup_items(t[None])
    np.testing.assert_allclose(it[0][0][0], t @ b @ t.T)
    np.testing.assert_allclose(it[0][1][0], t @ b @ t.conj().T)
    np.testing.assert_allclose(it[1][0][0], t.conj() @ b @ t.T)
    np.testing.assert_allclose(it[1][1][0], t.conj() @ b @ t.conj().T)
    np.testing.assert_allclose(it[0][2][0], t @ b)
    np.testing.assert_allclose(it[2][0][0], b @ t.T)


@pytest.mark.parametrize("kind", ["CompRxRx", "CompRyRy", "CompRzRz"])
def test_closed_form_sptm_equals_trace_formula(kind):
    """tests/test_operations/test_sptms/test_sptm_comp_rxrx.py:27-41 and siblings."""
    rng = np.random.default_rng(1)
    p = rng.uniform(0, 2 * np.pi, (7, 2))
    closed = Op(kind, (0, 1), params=p).to_sptm().sptm_matrix()
    generic = gates.enforce_real(gates.sptm_from_gate(Op(kind, (0, 1), params=p).matchgate_matrix()))
    np.testing.assert_allclose(closed, generic, atol=1e-14)
    fs = gates.enforce_real(gates.sptm_from_gate(gates.matchgate_fswap()))
    np.testing.assert_allclose(fs, gates.sptm_fswap((0, 1)), atol=1e-15)


@pytest.mark.parametrize("n", [3, 5, 8])
def test_fswap_chain_equals_long_range_fswap(n):
    """tests/test_operations/test_sptms/test_sptm_fswap.py:61-88."""
    ops = [Op("SptmFSwap", (i, i + 1)) for i in range(n - 1)]
    r = gates.chain(ops, n, contraction=None)
    long = gates.pad(gates.sptm_fswap((0, n - 1)), tuple(range(n)), tuple(range(n)))
    # a chain of adjacent fSWAPs moves wire 0 to the end: same permutation as the transposed long-range block
    assert np.allclose(r @ r.T, np.eye(2 * n))
    assert set(np.unique(r)) <= {0.0, 1.0}
    assert np.allclose(long @ long.T, np.eye(2 * n))


def test_product_state_equals_lookup_table():
    """tests/test_devices/test_probability_strategies/test_product_state_strategy.py:112-147."""
    rng = np.random.default_rng(5)
    n = 5
    r = gates.random_sptm(n, seed=3)
    bits = rng.integers(0, 2, n)
    amp = cv.amplitudes_from_basis_state(bits)
    states = readout.states_to_binary(np.arange(2 ** 3), 3)
    w = np.array([4, 0, 2])
    p_lt = readout.probs_lookup_table(r, bits, states, w)
    p_ps = readout.probs_product_state(cv.evolve(cv.covariance_matrix(amp), r), states, w)
    np.testing.assert_allclose(p_lt, p_ps, atol=1e-13)
    assert abs(p_ps.sum() - 1) < 1e-12


def test_gram_index_rules():
    """tests/test_ml/test_kernels/test_gram_matrix.py:40-68,84-122."""
    for shape in [(10, 10), (12, 10), (10, 12), (4, 5), (5, 4)]:
        idx = np.concatenate(list(kernels.indices_batch_generator(shape, 7)))
        exp = np.stack(np.triu_indices(shape[0], 1, shape[1]) if shape[0] < shape[1]
                       else np.tril_indices(shape[0], -1, shape[1]), axis=1)
        assert sorted(map(tuple, idx)) == sorted(map(tuple, exp))
        g = kernels.GramMatrixOracle(shape).apply_(lambda ids: -1.0)
        for i, j in np.ndindex(shape):
            assert g.data[i, j] == (1.0 if i == j else -1.0)


def test_kernel_properties_and_iris():
    """test_fermionic_pqc_kernel.py:111-120 (identical rows => ones, symmetry) + iris accuracy."""
    from sklearn import datasets
    from sklearn.model_selection import train_test_split
    from sklearn.preprocessing import MinMaxScaler
    from sklearn.svm import SVC

    k = kernels.FermionicPQCKernelOracle(n_qubits=4, rotations="X,Z").fit(np.ones((5, 8)))
    np.testing.assert_allclose(k.gram(np.ones((5, 8)) * 0.3), 1.0, atol=1e-6)
    gi = G["iris_accuracy"]
    x, y = datasets.load_iris(return_X_y=True)
    xtr, xte, ytr, yte = train_test_split(x, y, test_size=gi["test_size"], random_state=gi["random_state"])
    sc = MinMaxScaler((0, 1)).fit(xtr)
    k = kernels.FermionicPQCKernelOracle(n_qubits=gi["n_qubits"], rotations=gi["rotations"]).fit(sc.transform(xtr))
    ktr = k.gram(sc.transform(xtr))
    np.testing.assert_array_equal(ktr, ktr.T)
    clf = SVC(kernel="precomputed").fit(ktr, ytr)
    assert clf.score(k.transform(sc.transform(xte)), yte) == pytest.approx(gi["value"])


def test_gram_identity_covariance_sum():
    """K_ij = 2^-N |Pf(Lambda_i + Lambda_j)| equals the reference's per-pair formulation."""
    rng = np.random.default_rng(0)
    n = 6
    k = kernels.FermionicPQCKernelOracle(n_qubits=n, rotations="X,Z").fit(rng.random((9, 2 * n)))
    x = rng.random((9, 2 * n))
    s = k.x_to_sptm(x)
    idx = np.stack(np.tril_indices(9, -1), axis=1)
    ref = k.pair_expvals(s, s, idx)
    lam0 = cv.covariance_matrix(cv.amplitudes_from_basis_state(np.zeros(n, int)))
    lam = cv.evolve(lam0, s)
    alt = 2.0 ** -n * np.abs(pfaffian.pfaffian_signed(lam[idx[:, 0]] + lam[idx[:, 1]]))
    np.testing.assert_allclose(ref, alt, atol=1e-13)


# ------------------------------------------------------------------ dense state-vector simulation
def _random_matchgate_circuit(rng, n, n_gates):
    """Closed-form rotations, fSWAPs and generic matchgates (a product of three rotations, which goes through the trace
    formula R = 1/4 Tr(U c_mu U^+ c_nu) of utils/__init__.py:561-596 instead of a closed form)."""
    ops = []
    for _ in range(n_gates):
        w = int(rng.integers(0, n - 1))
        kind = ("CompRxRx", "CompRyRy", "CompRzRz", "fSWAP", "Matchgate")[int(rng.integers(0, 5))]
        if kind == "Matchgate":
            u = gates.matchgate_comp_rotation(rng.uniform(0, 2 * np.pi, 2), "XX") @ \
                gates.matchgate_comp_rotation(rng.uniform(0, 2 * np.pi, 2), "ZZ") @ \
                gates.matchgate_comp_rotation(rng.uniform(0, 2 * np.pi, 2), "YY")
            ops.append(gates.Op(kind, (w, w + 1), matrix=u))
        else:
            ops.append(gates.Op(kind, (w, w + 1), params=None if kind == "fSWAP" else rng.uniform(0, 2 * np.pi, 2)))
    return ops


@pytest.mark.parametrize("n", [2, 3, 4, 5, 6, 7])
def test_oracle_probabilities_equal_dense_statevector(n):
    """The reference pins its device against qml.device("default.qubit") on small circuits
    (tests/test_devices/test_nif_device/test_multiples_matchgate_circuit.py:66-102,
    test_product_state_strategy.py:50-61).  Same check for the oracle, with a hand-written 2^n state-vector simulation
    of the complex matchgates (tests/helpers.py): full distributions, marginals (in both wire orders) and single
    outcomes, from random basis states."""
    from tests.helpers import statevector_probs

    rng = np.random.default_rng(100 + n)
    for _ in range(3):
        ops = _random_matchgate_circuit(rng, n, 4 * n)
        init = rng.integers(0, 2, n)
        r = gates.chain(ops, n)
        amp = cv.amplitudes_from_basis_state(init)
        for wires in (list(range(n)), [0], [n - 1], [n - 1, 0], list(range(n))[::2]):
            ref = statevector_probs(ops, n, init, wires)
            got = readout.analytic_probability(r, amp, wires).reshape(-1)
            np.testing.assert_allclose(got, ref, rtol=0, atol=5e-14)
        # the two probability strategies of the device on one outcome (lookup table / product-state formula)
        tgt = rng.integers(0, 2, n)
        ref = statevector_probs(ops, n, init)[int("".join(map(str, tgt)), 2)]
        np.testing.assert_allclose(readout.probs_lookup_table(r, init, tgt, np.arange(n)), ref, rtol=0, atol=5e-14)
        cov_t = cv.evolve(cv.covariance_matrix(amp), r)
        np.testing.assert_allclose(readout.probs_product_state(cov_t, tgt, np.arange(n)), ref, rtol=0, atol=5e-14)


@pytest.mark.parametrize("n,rotations,nf", [(4, "X,Z", 8), (4, "Y,Z", 8), (6, "X,Y,Z", 6), (4, "Z", 4)])
def test_oracle_kernel_equals_dense_statevector_kernel(n, rotations, nf):
    """tests/test_ml/test_kernels/test_fermionic_pqc_kernel.py:122-129: the fermionic kernel equals the fidelity kernel
    |<psi(x_j)|psi(x_i)>|^2 of the same ansatz simulated on a dense state vector."""
    from tests.helpers import statevector

    rng = np.random.default_rng(n * 7 + nf)
    x = rng.random((5, nf))
    k = kernels.FermionicPQCKernelOracle(n_qubits=n, rotations=rotations).fit(x)
    got = k.gram(x, float32_store=False)

    def unbatched(ops, i):
        return [gates.Op(o.kind, o.wires, params=None if o.params is None else np.asarray(o.params)[i]) for o in ops]

    ops = k.ansatz(x)
    states = [statevector(unbatched(ops, i), n).reshape(-1) for i in range(len(x))]
    ref = np.array([[abs(np.vdot(a, b)) ** 2 for b in states] for a in states])
    np.testing.assert_allclose(got, ref, rtol=0, atol=5e-14)


@pytest.mark.parametrize("n", [2, 3, 4, 5])
def test_oracle_pauli_expvals_equal_dense_statevector(n):
    """qml.expval(sum of Pauli words) through the extended covariance and the sector Pfaffians
    (m_pfaffian_expval_strategy.py:270-306, jordan_wigner.py:62-140) against <psi| P |psi> on the dense state vector:
    every single-qubit letter, odd and even Majorana rank, identity, from random product states whose amplitudes are
    propagated through a random matchgate circuit (tests/test_devices/test_expval_strategies/test_m_pfaffian/
    test_m_pfaffian_expval_strategy.py:683-915 does this against default.qubit)."""
    from tests.helpers import statevector

    paulis = {"I": np.eye(2), "X": np.array([[0, 1], [1, 0]]), "Y": np.array([[0, -1j], [1j, 0]]),
              "Z": np.array([[1, 0], [0, -1]])}
    rng = np.random.default_rng(500 + n)
    for trial in range(3):
        ops = _random_matchgate_circuit(rng, n, 3 * n)
        r = gates.chain(ops, n)
        if trial == 0:
            amp = cv.amplitudes_from_basis_state(rng.integers(0, 2, n))
        else:
            amp = rng.standard_normal((n, 2)) + 1j * rng.standard_normal((n, 2))
            amp /= np.linalg.norm(amp, axis=1, keepdims=True)
        # product state |phi_0> ... |phi_{n-1}> as the initial dense vector, then the circuit
        psi0 = np.ones(1, dtype=complex)
        for q in range(n):
            psi0 = np.kron(psi0, amp[q])
        psi = np.zeros(2 ** n, dtype=complex)
        for b in range(2 ** n):
            bits = [(b >> (n - 1 - q)) & 1 for q in range(n)]
            psi += psi0[b] * statevector(ops, n, bits).reshape(-1)
        words = ["".join(rng.choice(list("IXYZ"), n)) for _ in range(10)] + ["I" * n, "Z" * n, "X" + "I" * (n - 1)]
        coeffs = rng.standard_normal(len(words))
        dense = 0.0
        for w, c in zip(words, coeffs):
            p = np.ones((1, 1), dtype=complex)
            for ch in w:
                p = np.kron(p, paulis[ch])
            dense += c * np.real(np.vdot(psi, p @ psi))
        got = readout.expval_pauli_sum(cv.evolved_extended_covariance(amp, r), words, coeffs)
        np.testing.assert_allclose(float(got), dense, rtol=0, atol=2e-13)
