"""GPU parity through the reference-facing Python API (device / kernels), read like the reference's own tests."""
import json
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")

import matchcake_b200 as mc
from matchcake_b200 import observables as obs
from matchcake_b200.ml.kernels import FermionicPQCKernel, Nystroem
from matchcake_b200.operations import (BasisState, CompRxRx, CompRyRy, CompRzRz, MatchgateOperation, ProductState,
                                       SingleParticleTransitionMatrixOperation, SptmCompRxRx, SptmCompRyRy,
                                       SptmCompRzRz, SptmFSwap, fSWAP)
from oracle import covariance as cv
from oracle import gates, kernels, readout
from oracle.gates import Op
from tests.helpers import readme_ops

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-10, 1e-13
G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_golden.json")))


def readme_circuit(params, wires, initial_state):
    """README.md:48-66 written against this package."""
    yield BasisState(initial_state, wires=wires)
    for even_wire in wires[:-1:2]:
        idx = list(wires).index(even_wire)
        curr = [wires[idx], wires[idx + 1]]
        yield CompRxRx(params, wires=curr)
        yield CompRyRy(params, wires=curr)
        yield CompRzRz(params, wires=curr)
    for odd_wire in wires[1:-1:2]:
        idx = list(wires).index(odd_wire)
        yield fSWAP(wires=[wires[idx], wires[idx + 1]])


def test_readme_circuit_joss_value():
    dev = mc.NonInteractingFermionicDevice(wires=4)
    init = np.zeros(4, dtype=int)
    v = dev.execute_generator(readme_circuit(np.array([0.1, 0.2]), dev.wires, init),
                              observable=obs.Projector(init, wires=dev.wires), output_type="expval", reset=True)
    assert f"{float(v):.4f}" == G["joss_expval"]["printed"]
    assert dev.apply_metadata["n_operations"] == 8
    assert dev.apply_metadata["n_contracted_operations"] == 2  # two layers under the neighbours strategy


@pytest.mark.parametrize("n,B", [(4, 17), (16, 128), (9, 3), (32, 5), (33, 4), (48, 3), (64, 2), (80, 2), (128, 2)])
def test_batched_projector_expval_and_sptm(n, B):
    rng = np.random.default_rng(n)
    p = rng.uniform(0, 2 * np.pi, (B, 2))
    init = np.zeros(n, dtype=int)
    dev = mc.NonInteractingFermionicDevice(wires=n)
    v = dev.execute_generator(readme_circuit(p, list(dev.wires), init), observable=obs.Projector(init, wires=dev.wires),
                              output_type="expval", reset=True)
    r_ref = gates.chain(readme_ops(n, p), n)
    ref = readout.probs_lookup_table(r_ref, init, init, np.arange(n))
    np.testing.assert_allclose(v.numpy(), ref, rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(dev.global_sptm.matrix().cpu().numpy(), r_ref, rtol=RTOL, atol=ATOL)
    # marginal probabilities on a wire subset, both strategies of the reference give the same numbers
    wires = [n - 1, 0, 2]
    pr = dev.analytic_probability(wires).numpy()
    np.testing.assert_allclose(pr, readout.analytic_probability(r_ref, cv.amplitudes_from_basis_state(init), wires),
                               rtol=RTOL, atol=ATOL)
    assert pr.shape == (B, 8)
    # a partial projector goes through the unfused path
    tgt = [1, 0]
    v2 = dev.get_states_probability(np.array(tgt), [1, 3]).numpy()
    np.testing.assert_allclose(v2, readout.probs_lookup_table(r_ref, init, tgt, np.array([1, 3])), rtol=RTOL, atol=ATOL)


def test_probabilities_tutorial_golden():
    wires = [0, 1, 2]
    sptm = SingleParticleTransitionMatrixOperation.random(wires=wires, seed=0)
    th = np.pi / 3
    mixed = np.array([[1.0, 1.0], [1.0, 0.0], [np.cos(th / 2), np.sin(th / 2)]], dtype=complex)
    mixed /= np.linalg.norm(mixed, axis=1, keepdims=True)
    preps = {"probs_basis_101": ProductState.from_basis_state([1, 0, 1], wires=wires),
             "probs_plus": ProductState(np.full((3, 2), 1 / np.sqrt(2), dtype=complex), wires=wires),
             "probs_mixed": ProductState(mixed, wires=wires)}
    for key, prep in preps.items():
        dev = mc.NIFDevice(wires=3)
        p = dev.execute_generator([prep, sptm], output_type="probs", reset=True).numpy()
        np.testing.assert_array_equal(np.round(p, 6) + 0.0, np.asarray(G[key]["value"]))
    dev = mc.NIFDevice(wires=3)
    dev.execute_generator([preps["probs_basis_101"], sptm], reset=True)
    np.testing.assert_array_equal(np.round(dev.probability(wires=[1]).numpy(), 6), G["probs_marginal_q1"]["value"])
    assert round(float(dev.expval(obs.Projector([1, 0, 1], wires=wires))), 6) == G["projector_101"]["value"]
    assert round(float(dev.get_states_probability("101")), 6) == G["projector_101"]["value"]
    assert round(float(dev.get_states_probability(5)), 6) == G["projector_101"]["value"]


@pytest.mark.parametrize("n", [2, 3, 5, 6, 8])
def test_device_probabilities_equal_dense_statevector(n):
    """The reference's own device tests: the same circuit on qml.device("default.qubit")
    (tests/test_devices/test_nif_device/test_multiples_matchgate_circuit.py:66-102).  Here the dense side is the
    hand-written 2^n state-vector simulation of tests/helpers.py: qml.probs over all wires and over marginals (both wire
    orders), and the projector expectation value, from a random basis state."""
    from tests.helpers import statevector_probs

    rng = np.random.default_rng(300 + n)
    classes = {"CompRxRx": CompRxRx, "CompRyRy": CompRyRy, "CompRzRz": CompRzRz}
    for _ in range(2):
        ref_ops, dev_ops = [], []
        init = rng.integers(0, 2, n)
        dev_ops.append(BasisState(init, wires=range(n)))
        for _ in range(4 * n):
            w = int(rng.integers(0, n - 1))
            kind = ("CompRxRx", "CompRyRy", "CompRzRz", "fSWAP")[int(rng.integers(0, 4))]
            if kind == "fSWAP":
                ref_ops.append(Op(kind, (w, w + 1)))
                dev_ops.append(fSWAP(wires=[w, w + 1]))
            else:
                p = rng.uniform(0, 2 * np.pi, 2)
                ref_ops.append(Op(kind, (w, w + 1), params=p))
                dev_ops.append(classes[kind](p, wires=[w, w + 1]))
        dev = mc.NonInteractingFermionicDevice(wires=n)
        dev.execute_generator(dev_ops, output_type=None, reset=True)
        for wires in (list(range(n)), [0], [n - 1], [n - 1, 0], list(range(n))[::2]):
            got = np.asarray(dev.analytic_probability(wires)).reshape(-1)
            np.testing.assert_allclose(got, statevector_probs(ref_ops, n, init, wires), rtol=0, atol=5e-14)
        tgt = rng.integers(0, 2, n)
        v = dev.execute_generator(dev_ops, observable=obs.Projector(tgt, wires=range(n)), output_type="expval", reset=True)
        ref = statevector_probs(ref_ops, n, init)[int("".join(map(str, tgt)), 2)]
        np.testing.assert_allclose(float(v), ref, rtol=0, atol=5e-14)


def test_expectation_values_tutorial_golden():
    g = G["hamiltonian"]
    h = obs.Hamiltonian(g["coeffs"], [obs.X(0) @ obs.X(1), obs.Y(1) @ obs.Y(2), obs.Z(0) @ obs.Z(1), obs.Y(1) @ obs.X(2),
                                      obs.X(0) @ obs.Y(1), obs.Z(1) @ obs.I(2), obs.I(1) @ obs.Z(2)])
    wires = [0, 1, 2]
    sptm = SingleParticleTransitionMatrixOperation.random(wires=wires, seed=0)
    th = np.pi / 3
    mixed = np.array([[1.0, 1.0], [1.0, 0.0], [np.cos(th / 2), np.sin(th / 2)]], dtype=complex)
    mixed /= np.linalg.norm(mixed, axis=1, keepdims=True)
    for key, prep in [("expval_basis_101", ProductState.from_basis_state([1, 0, 1], wires=wires)),
                      ("expval_plus", ProductState(np.full((3, 2), 1 / np.sqrt(2), dtype=complex), wires=wires)),
                      ("expval_mixed", ProductState(mixed, wires=wires))]:
        dev = mc.NIFDevice(3)
        e = dev.execute_generator([prep, sptm], observable=h, output_type="expval", reset=True)
        assert round(float(e), 6) == g[key]


def test_mixed_stream_dense_and_gates_batched():
    """Gate segments interleaved with dense SPTM ops on sub-ranges, batched, generic matchgates, product-state input."""
    rng = np.random.default_rng(4)
    n, B = 6, 5
    p1, p2 = rng.uniform(0, 6, (B, 2)), rng.uniform(0, 6, (2,))
    dense = gates.random_sptm(3, seed=5, batch_size=B)
    mg = gates.matchgate_comp_rotation(rng.uniform(0, 6, (2,)), "YY") @ gates.matchgate_comp_rotation(rng.uniform(0, 6, (2,)), "XX")
    stream = [SptmCompRyRy(p1, wires=[0, 1]), SptmCompRzRz(p2, wires=[2, 3]), SptmFSwap(wires=[1, 4]),
              SingleParticleTransitionMatrixOperation(dense, wires=[2, 3, 4]), MatchgateOperation(mg, wires=[4, 5]),
              SptmCompRxRx(p1, wires=[3, 4]), SptmFSwap(wires=[5, 3])]
    oops = [Op("SptmCompRyRy", (0, 1), params=p1), Op("SptmCompRzRz", (2, 3), params=p2), Op("SptmFSwap", (1, 4)),
            Op("Sptm", (2, 3, 4), matrix=dense), Op("Matchgate", (4, 5), matrix=mg), Op("SptmCompRxRx", (3, 4), params=p1),
            Op("SptmFSwap", (5, 3))]
    r_ref = gates.chain(oops, n)
    amp = rng.standard_normal((n, 2)) + 1j * rng.standard_normal((n, 2))
    amp /= np.linalg.norm(amp, axis=1, keepdims=True)
    dev = mc.NIFDevice(n)
    dev.execute_generator([ProductState(amp, wires=range(n))] + stream, reset=True)
    np.testing.assert_allclose(dev.global_sptm.matrix().cpu().numpy(), r_ref, rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(dev.covariance_matrix.numpy(), cv.evolve(cv.covariance_matrix(amp), r_ref), rtol=RTOL, atol=ATOL)
    pr = dev.analytic_probability([0, 3, 5]).numpy()
    np.testing.assert_allclose(pr, readout.analytic_probability(r_ref, amp, [0, 3, 5]), rtol=RTOL, atol=ATOL)
    words = ["".join(rng.choice(list("IXYZ"), n)) for _ in range(12)]
    coeffs = rng.standard_normal(12)
    h = obs.Hamiltonian(list(coeffs), [obs.PauliWord({i: c for i, c in enumerate(w)}) for w in words])
    np.testing.assert_allclose(dev.expval(h).numpy(), readout.expval_pauli_sum(cv.evolved_extended_covariance(amp, r_ref), words, coeffs),
                               rtol=RTOL, atol=1e-12)
    # per-outcome wires
    tg = np.array([[1, 0], [0, 1], [1, 1]])
    ws = np.array([[0, 1], [2, 5], [0, 1]])
    got = dev.get_states_probability(tg, ws).numpy()
    ref = np.stack([readout.probs_product_state(cv.evolve(cv.covariance_matrix(amp), r_ref), tg[i], ws[i]) for i in range(3)])
    np.testing.assert_allclose(got, ref, rtol=RTOL, atol=ATOL)


def test_device_error_behaviour():
    dev = mc.NIFDevice(3)
    with pytest.raises(mc.DeviceError):
        dev.execute_generator([SptmFSwap(wires=[0, 1]), BasisState([0, 0, 0], wires=[0, 1, 2])], reset=True)

    class Bad:
        wires = (0, 1)

        def matrix(self):
            return np.ones((4, 4))

    with pytest.raises(mc.DeviceError):
        dev.execute_generator([Bad()], reset=True)
    dev.reset()
    with pytest.raises(ValueError):
        dev.execute_generator([SptmFSwap(wires=[0, 1])], output_type="nope", reset=True)
    with pytest.raises(ValueError):
        dev.get_states_probability([0, 2, 1])
    np.testing.assert_array_equal(mc.NIFDevice(3).global_sptm.matrix().cpu().numpy(), np.eye(6))


# ------------------------------------------------------------------ kernels
def test_iris_pipeline_accuracy():
    """README.md:99-118 / tutorials/iris_classification.ipynb:1699."""
    from sklearn import datasets
    from sklearn.model_selection import train_test_split
    from sklearn.pipeline import Pipeline
    from sklearn.preprocessing import MinMaxScaler
    from sklearn.svm import SVC

    x, y = datasets.load_iris(return_X_y=True)
    xtr, xte, ytr, yte = train_test_split(x, y, test_size=0.2, random_state=0)
    pipe = Pipeline([("scaler", MinMaxScaler(feature_range=(0, 1))),
                     ("kernel", FermionicPQCKernel(n_qubits=4, rotations="X,Z").freeze()),
                     ("classifier", SVC(kernel="precomputed"))])
    pipe.fit(xtr, ytr)
    assert pipe.score(xte, yte) == pytest.approx(G["iris_accuracy"]["value"])


@pytest.mark.parametrize("n,nf,rot,ent", [(4, 4, "X,Z", "fswap"), (4, 7, "Y,Z", "fswap"), (8, 16, "X", "identity"),
                                          (6, 15, "Z,X", "fswap"), (32, 64, "X,Z", "fswap")])
def test_kernel_gram_equals_reference_formulation(n, nf, rot, ent):
    rng = np.random.default_rng(n + nf)
    x0, x1 = rng.random((12, nf)), rng.random((7, nf))
    k = FermionicPQCKernel(n_qubits=n, rotations=rot, entangling_mth=ent, r_dtype=torch.float64, c_dtype=torch.complex128)
    k.float32_gram = False
    k.fit(x0)
    o = kernels.FermionicPQCKernelOracle(n_qubits=n, rotations=rot, entangling_mth=ent).fit(x0)
    np.testing.assert_array_equal(k.bias_.detach().cpu().numpy(), o.bias_)
    assert k.depth_ == o.depth_
    np.testing.assert_allclose(k._x_to_sptm(x0).cpu().numpy(), o.x_to_sptm(x0), rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(k(x0).cpu().numpy(), o.gram(x0, float32_store=False), rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(k(x0, x1).cpu().numpy(), o.gram(x0, x1, float32_store=False), rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(k(x1, x0).cpu().numpy(), o.gram(x1, x0, float32_store=False), rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(k.transform(x1), o.transform(x1, float32_store=False), rtol=1e-6)
    # op-object view of the ansatz == compiled view
    dev = k.q_device
    dev.execute_generator(k.ansatz(x0), reset=True)
    np.testing.assert_allclose(dev.global_sptm.matrix().cpu().numpy(), o.x_to_sptm(x0), rtol=RTOL, atol=ATOL)
    # float32 default storage
    k32 = FermionicPQCKernel(n_qubits=n, rotations=rot, entangling_mth=ent).fit(x0)
    g32 = k32(x0)
    assert g32.dtype == torch.float32
    np.testing.assert_array_equal(g32.cpu().numpy(), o.gram(x0))


def test_kernel_identical_rows_and_symmetry():
    """tests/test_ml/test_kernels/test_fermionic_pqc_kernel.py:111-120."""
    x = np.full((6, 8), 0.25)
    k = FermionicPQCKernel(n_qubits=4).fit(x)
    np.testing.assert_allclose(k(x).cpu().numpy(), 1.0, atol=1e-6)
    rng = np.random.default_rng(0)
    x = rng.random((20, 8))
    g = k(x).cpu().numpy()
    np.testing.assert_array_equal(g, g.T)
    np.testing.assert_array_equal(np.diag(g), 1.0)


def test_nystroem_matches_oracle():
    """tutorials/nystroem_kernel_approximation.ipynb pipeline on synthetic data."""
    rng = np.random.default_rng(1)
    x = rng.random((40, 12))
    ny = Nystroem(FermionicPQCKernel(n_qubits=6, rotations="X,Z"), n_components=10, random_state=0).fit(x)
    ref = kernels.NystroemOracle(kernels.FermionicPQCKernelOracle(n_qubits=6, rotations="X,Z"), n_components=10, random_state=0).fit(x)
    np.testing.assert_array_equal(ny.component_indices_, ref.component_indices_)
    np.testing.assert_allclose(ny.normalization_, ref.normalization_, rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(ny.transform(x), ref.transform(x), rtol=1e-4, atol=1e-5)


def test_utils_pfaffian_namespace():
    """tests/test_utils/test_pfaffian.py of the reference, through matchcake_b200.utils."""
    from matchcake_b200 import utils

    m = np.array([[0.0, 3.0], [-3.0, 0.0]])
    out = utils.pfaffian(m, sign=True)
    assert isinstance(out, np.ndarray) and float(out) == 3.0
    assert float(utils.signed_pfaffian(torch.zeros(0, 0, dtype=torch.float64))) == 1.0
    assert float(utils.signed_pfaffian(torch.zeros(3, 3, dtype=torch.float64))) == 0.0
    t32 = torch.randn(4, 4)
    assert utils.signed_pfaffian(t32 - t32.T).dtype == torch.float32
    k = G["pfaffian_kats"]["sector_2x2"]
    np.testing.assert_allclose(utils.sector_pfaffian_features(np.array(k["cov"]), np.array(k["index_sets"])), k["value"])
    rng = np.random.default_rng(0)
    a = rng.random((3, 4, 4))
    a = a - np.swapaxes(a, -1, -2)
    np.testing.assert_allclose(utils.pfaffian(a) ** 2, np.abs(np.linalg.det(a)), rtol=1e-10)
    # epsilon is honoured like the reference's PfaffianDet: sqrt(|det A| + epsilon) (utils/_pfaffian.py:110-118)
    from oracle import pfaffian as opf
    tiny = 1e-7 * a
    for eps in (1e-32, 1e-20, 0.0):
        np.testing.assert_allclose(utils.pfaffian(tiny, epsilon=eps), opf.pfaffian_det(tiny, eps), rtol=1e-10, atol=0)
    assert np.all(utils.pfaffian(np.zeros((2, 4, 4))) == 1e-16)          # exact zeros come back as sqrt(1e-32)
    assert np.all(utils.pfaffian(np.zeros((2, 4, 4)), epsilon=0.0) == 0.0)
    big = rng.random((2, 70, 70))
    big = 1e-3 * (big - np.swapaxes(big, -1, -2))                          # block-tile kernel path
    np.testing.assert_allclose(utils.pfaffian(big, epsilon=1e-120), opf.pfaffian_det(big, 1e-120), rtol=1e-9)
    np.testing.assert_allclose(utils.pfaffian(big, epsilon=1e-2), opf.pfaffian_det(big, 1e-2), rtol=1e-10)


@pytest.mark.parametrize("n,nf,rot,ent", [(8, 8, "X,Z", "fswap"), (6, 5, "Y,Z", "fswap"), (10, 10, "X", "identity"),
                                          (4, 4, "Z,Y,X", "hadamard")])
def test_block_factorised_gram_equals_dense_and_oracle(n, nf, rot, ent):
    """depth_ = 1 ansatz: the SPTM is block diagonal and the Gram factorises into 2-qubit factors (gram_blocks.cu)."""
    rng = np.random.default_rng(n * 7 + nf)
    x0, x1 = rng.random((70, nf)), rng.random((9, nf))
    k = FermionicPQCKernel(n_qubits=n, rotations=rot, entangling_mth=ent, r_dtype=torch.float64, c_dtype=torch.complex128)
    k.float32_gram = False
    k.fit(x0)
    assert k.depth_ == 1 and k._block_plan() is not None
    fast = [k(x0).cpu().numpy(), k(x0, x1).cpu().numpy(), k(x1, x0).cpu().numpy(), k.transform(x1)]
    k.use_block_factorisation = False
    dense = [k(x0).cpu().numpy(), k(x0, x1).cpu().numpy(), k(x1, x0).cpu().numpy(), k.transform(x1)]
    for a, b in zip(fast, dense):
        np.testing.assert_allclose(a, b, rtol=RTOL, atol=ATOL)
    if ent != "hadamard":
        o = kernels.FermionicPQCKernelOracle(n_qubits=n, rotations=rot, entangling_mth=ent).fit(x0)
        np.testing.assert_allclose(fast[0], o.gram(x0, float32_store=False), rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(fast[1], o.gram(x0, x1, float32_store=False), rtol=RTOL, atol=ATOL)
    # a depth-2 ansatz couples the pairs: no factorisation
    k2 = FermionicPQCKernel(n_qubits=4, rotations="X,Z").fit(rng.random((5, 8)))
    assert k2.depth_ == 2 and k2._block_plan() is None


def test_plain_c_client_of_the_abi(tmp_path):
    """tests/c/abi_client.c: a C99 program (no Python, no torch) drives the library through include/mcb200.h --
    Pfaffian KATs, the README circuit (0.9901) fused and as the three-call sequence, error reporting."""
    import subprocess

    from tests.test_library_abi import _compile_c_client

    exe = _compile_c_client(str(tmp_path / "abi_client"))
    res = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "abi_client ok" in res.stdout


def test_strategy_objects_drive_the_gpu_path():
    """matchcake_b200.strategies: the reference's plug-in signatures (probability / expval / contraction strategy)
    called the way its device calls them, against the oracle."""
    from matchcake_b200 import strategies as st
    from matchcake_b200.operations import BasisState, ProductState

    n, B = 4, 3
    rng = np.random.default_rng(21)
    r = gates.random_sptm(n, seed=2, batch_size=B)
    bits = np.array([1, 0, 0, 1])
    targets = np.array([[0, 1], [1, 1], [0, 0]])
    wires = [2, 0]
    p = st.B200ProbabilityStrategy()(state_prep_op=BasisState(bits, wires=range(n)), target_binary_states=targets,
                                     wires=wires, global_sptm=r, all_wires=list(range(n)))
    ref = readout.states_probability(r, cv.amplitudes_from_basis_state(bits), targets, np.array(wires))
    np.testing.assert_allclose(p.cpu().numpy(), ref, rtol=RTOL, atol=ATOL)
    amp = rng.standard_normal((n, 2)) + 1j * rng.standard_normal((n, 2))
    amp /= np.linalg.norm(amp, axis=1, keepdims=True)
    words, coeffs = ["ZIII", "IXXI", "YZZX"], rng.standard_normal(3)
    h = obs.Hamiltonian(list(coeffs), [obs.PauliWord({i: c for i, c in enumerate(w)}) for w in words])
    e = st.B200PfaffianExpvalStrategy()(ProductState(amp, wires=range(n)), h, global_sptm=r, all_wires=list(range(n)))
    np.testing.assert_allclose(e.cpu().numpy(), readout.expval_pauli_sum(cv.evolved_extended_covariance(amp, r), words, coeffs),
                               rtol=RTOL, atol=1e-12)
    pr = rng.uniform(0, 6, (B, 2))
    stream = [SptmCompRyRy(pr, wires=[0, 1]), SptmFSwap(wires=[1, 3]), SptmCompRxRx(pr, wires=[2, 3])]
    out = st.B200ContractionStrategy(wires=range(n))(stream)
    assert len(out) == 1
    ref_r = gates.chain([Op("SptmCompRyRy", (0, 1), params=pr), Op("SptmFSwap", (1, 3)), Op("SptmCompRxRx", (2, 3), params=pr)], n)
    np.testing.assert_allclose(out[0].matrix().cpu().numpy(), ref_r, rtol=RTOL, atol=ATOL)
