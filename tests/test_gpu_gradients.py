"""Adjoint kernels (csrc/backward.cu) against reference gradients.

The reference differentiates its path with torch autograd, so the checks are: (i) torch autograd over the torch-CPU
restatement of the reference op sequence (oracle/ref_torch.py), (ii) central finite differences of the numpy oracle,
(iii) the closed form d Pf = Pf/2 tr(A^-1 dA) that torch_pfaffian implements (reference tests
tests/test_utils/test_pfaffian.py:86-115 run the same gradchecks).  Tolerances: analytic comparisons rtol 1e-9,
finite differences rtol 2e-6 (h = 1e-6, float64).
"""
import numpy as np
import pytest
import torch

from oracle import covariance as cv
from oracle import gates, readout, ref_torch
from oracle.gates import Op
from tests.helpers import random_skew

pytestmark = pytest.mark.gpu

import matchcake_b200 as mc  # noqa: E402
from matchcake_b200 import observables as obs  # noqa: E402
from matchcake_b200.operations import (BasisState, CompRxRx, CompRyRy, CompRzRz, SingleParticleTransitionMatrixOperation,  # noqa: E402
                                       SptmCompRxRx, SptmCompRyRy, SptmCompRzRz, SptmFSwap, fSWAP)


def dev(a, grad=False):
    t = torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device="cuda")
    return t.requires_grad_() if grad else t


def finite_difference(f, x, h=1e-6):
    """Central differences of a scalar function of a numpy array."""
    g = np.zeros_like(x)
    it = np.nditer(x, flags=["multi_index"])
    for _ in it:
        i = it.multi_index
        xp, xm = x.copy(), x.copy()
        xp[i] += h
        xm[i] -= h
        g[i] = (f(xp) - f(xm)) / (2 * h)
    return g


# ------------------------------------------------------------------ K3
@pytest.mark.parametrize("m", [2, 4, 6, 8, 16, 30, 32, 48, 64, 70, 128])
@pytest.mark.parametrize("signed", [False, True])
def test_pfaffian_backward_closed_form(m, signed):
    from matchcake_b200 import ops as mops

    n = 5
    a_np = random_skew(np.random.default_rng(m), (n, m, m))
    a = dev(a_np, grad=True)
    pf = mops.pfaffian(a, signed=signed)
    w = np.random.default_rng(m + 1).standard_normal(n)
    (pf * dev(w)).sum().backward()
    pf_np = pf.detach().cpu().numpy()
    expect = (w * pf_np / 2)[:, None, None] * np.swapaxes(np.linalg.inv(a_np), 1, 2)
    scale = np.abs(expect).max()
    np.testing.assert_allclose(a.grad.cpu().numpy(), expect, rtol=1e-8, atol=1e-9 * scale)


def test_pfaffian_gradcheck_like_reference():
    """tests/test_utils/test_pfaffian.py:86-115: magnitude, zeros, signed on the skew manifold."""
    from matchcake_b200.utils import pfaffian

    a = dev(random_skew(np.random.default_rng(0), (4, 4)), grad=True)
    assert torch.autograd.gradcheck(lambda x: pfaffian(x, sign=False), (a,), atol=1e-5, rtol=1e-4)
    z = torch.zeros(4, 4, dtype=torch.float64, device="cuda", requires_grad=True)
    assert torch.autograd.gradcheck(lambda x: pfaffian(x, sign=False), (z,), atol=1e-5, rtol=1e-4)
    n = 6
    iu = torch.triu_indices(n, n, 1, device="cuda")

    def skew_from_upper(t):
        m = torch.zeros(n, n, dtype=torch.float64, device="cuda")
        m = m.index_put((iu[0], iu[1]), t)
        return m - m.T

    theta = torch.randn(n * (n - 1) // 2, dtype=torch.float64, device="cuda", requires_grad=True)
    assert torch.autograd.gradcheck(lambda t: pfaffian(skew_from_upper(t), sign=True), (theta,), atol=1e-5, rtol=1e-4)


def test_singular_matrix_gives_zero_gradient():
    from matchcake_b200 import ops as mops

    a = torch.zeros(3, 6, 6, dtype=torch.float64, device="cuda")
    a[0, 0, 1], a[0, 1, 0] = 1.0, -1.0  # rank 2: Pf = 0
    a.requires_grad_()
    mops.pfaffian(a, signed=True).sum().backward()
    assert torch.all(a.grad == 0)


# ------------------------------------------------------------------ K1 + K2 + K3 through the device
@pytest.mark.parametrize("n", [4, 6, 10])
def test_readme_circuit_gradients_match_reference_autograd(n):
    """d expval / d params of the README circuit: adjoint kernels vs torch autograd over the reference op sequence."""
    B = 5
    p_np = np.random.default_rng(n).uniform(0, 2 * np.pi, (B, 2))
    w_np = np.random.default_rng(n + 1).standard_normal(B)
    pr = torch.tensor(p_np, dtype=torch.float64, requires_grad=True)
    ref = ref_torch.projector_zero_lookup(ref_torch.readme_circuit_sptm(pr, n))
    (ref * torch.as_tensor(w_np)).sum().backward()

    p = dev(p_np, grad=True)
    d = mc.NIFDevice(n)
    ops_ = [BasisState(np.zeros(n, dtype=int), wires=range(n))]
    for w in range(0, n - 1, 2):
        ops_ += [CompRxRx(p, wires=[w, w + 1]), CompRyRy(p, wires=[w, w + 1]), CompRzRz(p, wires=[w, w + 1])]
    for w in range(1, n - 1, 2):
        ops_.append(fSWAP(wires=[w, w + 1]))
    e = d.execute_generator(ops_, observable=obs.Projector(np.zeros(n, dtype=int), wires=range(n)),
                            output_type="expval", reset=True)
    np.testing.assert_allclose(e.detach().cpu().numpy(), ref.detach().numpy(), rtol=1e-10, atol=1e-13)
    (e * dev(w_np).to(e.device)).sum().backward()
    np.testing.assert_allclose(p.grad.cpu().numpy(), pr.grad.numpy(), rtol=1e-8, atol=1e-11)


def _random_stream(n, B, rng):
    """Rotation gates with independent (B, 2) parameters and fSWAPs (adjacent and long range)."""
    kinds = [("SptmCompRxRx", SptmCompRxRx), ("SptmCompRyRy", SptmCompRyRy), ("SptmCompRzRz", SptmCompRzRz)]
    layout = []
    for layer in range(3):
        for w in range(layer % 2, n - 1, 2):
            layout.append((kinds[(layer + w) % 3], (w, w + 1)))
        layout.append((("SptmFSwap", SptmFSwap), (layer, min(n - 1, layer + 2))))
    n_par = sum(1 for (name, _), _ in layout if name != "SptmFSwap")
    params = rng.uniform(0, 2 * np.pi, (n_par, B, 2))
    return layout, params


def _oracle_ops(layout, params):
    out, j = [], 0
    for (name, _), wires in layout:
        if name == "SptmFSwap":
            out.append(Op(name, wires))
        else:
            out.append(Op(name, wires, params=params[j]))
            j += 1
    return out


def _device_ops(layout, ptens):
    out, j = [], 0
    for (name, cls), wires in layout:
        if name == "SptmFSwap":
            out.append(cls(wires=list(wires)))
        else:
            out.append(cls(ptens[j], wires=list(wires)))
            j += 1
    return out


@pytest.mark.parametrize("n,wires", [(5, [0, 2, 3]), (6, [1, 4])])
def test_probability_gradients_vs_finite_differences(n, wires):
    rng = np.random.default_rng(n)
    B = 2
    layout, params = _random_stream(n, B, rng)
    bits = rng.integers(0, 2, n)
    amp = cv.amplitudes_from_basis_state(bits)
    wq = rng.standard_normal((B, 2 ** len(wires)))

    def f(pp):
        r = gates.chain(_oracle_ops(layout, pp), n)
        return float((readout.analytic_probability(r, amp, wires) * wq).sum())

    fd = finite_difference(f, params)
    ptens = [dev(params[j], grad=True) for j in range(len(params))]
    d = mc.NIFDevice(n)
    d.execute_generator([BasisState(bits, wires=range(n))] + _device_ops(layout, ptens), reset=True)
    pr = d.analytic_probability(wires)
    np.testing.assert_allclose(pr.detach().cpu().numpy(), readout.analytic_probability(gates.chain(_oracle_ops(layout, params), n), amp, wires),
                               rtol=1e-10, atol=1e-13)
    (pr * dev(wq).to(pr.device)).sum().backward()
    got = np.stack([t.grad.cpu().numpy() for t in ptens])
    np.testing.assert_allclose(got, fd, rtol=2e-6, atol=2e-8)


def test_pauli_sum_expval_gradients_vs_finite_differences():
    n, B = 4, 3
    rng = np.random.default_rng(7)
    layout, params = _random_stream(n, B, rng)
    bits = np.array([1, 0, 0, 1])
    amp = cv.amplitudes_from_basis_state(bits)
    words = ["ZIII", "IZZI", "XXII", "IYYI", "XZZX", "IIXY", "YZXI", "IIII"]
    coeffs = rng.standard_normal(len(words))
    wq = rng.standard_normal(B)

    def f(pp):
        r = gates.chain(_oracle_ops(layout, pp), n)
        return float((readout.expval_pauli_sum(cv.evolved_extended_covariance(amp, r), words, coeffs) * wq).sum())

    fd = finite_difference(f, params)
    ptens = [dev(params[j], grad=True) for j in range(len(params))]
    d = mc.NIFDevice(n)
    h = obs.Hamiltonian(list(coeffs), [obs.PauliWord({i: c for i, c in enumerate(w)}) for w in words])
    e = d.execute_generator([BasisState(bits, wires=range(n))] + _device_ops(layout, ptens), observable=h,
                            output_type="expval", reset=True)
    (e * dev(wq).to(e.device)).sum().backward()
    got = np.stack([t.grad.cpu().numpy() for t in ptens])
    np.testing.assert_allclose(got, fd, rtol=2e-6, atol=2e-8)


def test_dense_sptm_factor_gradients():
    """A dense SPTM factor in the stream: the chain step's adjoint is two more products on the DMMA kernel."""
    from matchcake_b200 import ops as mops

    rng = np.random.default_rng(3)
    for ta in (False, True):
        for tb in (False, True):
            a_np, b_np = rng.standard_normal((3, 6, 6)), rng.standard_normal((6, 6))
            a, b = dev(a_np, grad=True), dev(b_np, grad=True)
            w = dev(rng.standard_normal((3, 6, 6)))
            (mops.sptm_matmul(a, b, ta, tb) * w).sum().backward()
            ar, br = torch.tensor(a_np, requires_grad=True), torch.tensor(b_np, requires_grad=True)
            ((ar.transpose(1, 2) if ta else ar) @ (br.T if tb else br) * w.cpu()).sum().backward()
            np.testing.assert_allclose(a.grad.cpu().numpy(), ar.grad.numpy(), rtol=1e-11, atol=1e-12)
            np.testing.assert_allclose(b.grad.cpu().numpy(), br.grad.numpy(), rtol=1e-11, atol=1e-12)


def test_two_wire_sptm_block_gradients():
    """Batched explicit 4x4 blocks (BLOCK4_PARAM gates) receive the derivative of their 16 entries.  Entry-wise
    perturbations leave SO(4), so the finite differences use the covariance formulation (the one the kernels implement;
    it agrees with the lookup-table formulation only on the orthogonal group) and the circuit is generic enough that no
    outcome has probability 0 (|Pf| is not differentiable there)."""
    n, B = 4, 2
    rng = np.random.default_rng(11)
    blk = np.stack([gates.sptm_comp_ryry(rng.uniform(0, 6, 2)) @ gates.sptm_comp_rxrx(rng.uniform(0, 6, 2))
                    for _ in range(B)])
    p = rng.uniform(0, 6, (4, B, 2))
    bits = np.array([0, 1, 0, 0])
    amp = cv.amplitudes_from_basis_state(bits)
    wq = rng.standard_normal((B, 4))
    states = readout.states_to_binary(np.arange(4), 2)

    def probs_of(m):
        r = gates.chain([Op("SptmCompRxRx", (0, 1), params=p[0]), Op("SptmCompRyRy", (2, 3), params=p[1]),
                         Op("Sptm", (1, 2), matrix=m), Op("SptmCompRxRx", (2, 3), params=p[2]),
                         Op("SptmFSwap", (0, 2)), Op("SptmCompRyRy", (0, 1), params=p[3])], n)
        pr = readout.probs_product_state(cv.evolve(cv.covariance_matrix(amp), r), states, np.array([1, 3])).T
        return pr / pr.sum(axis=-1, keepdims=True)

    assert probs_of(blk).min() > 1e-3
    fd = finite_difference(lambda m: float((probs_of(m) * wq).sum()), blk)
    m = dev(blk, grad=True)
    d = mc.NIFDevice(n)
    d.execute_generator([BasisState(bits, wires=range(n)), SptmCompRxRx(dev(p[0]), wires=[0, 1]),
                         SptmCompRyRy(dev(p[1]), wires=[2, 3]), SingleParticleTransitionMatrixOperation(m, wires=[1, 2]),
                         SptmCompRxRx(dev(p[2]), wires=[2, 3]), SptmFSwap(wires=[0, 2]),
                         SptmCompRyRy(dev(p[3]), wires=[0, 1])], reset=True)
    pr = d.analytic_probability([1, 3])
    np.testing.assert_allclose(pr.detach().cpu().numpy(), probs_of(blk), rtol=1e-10, atol=1e-13)
    (pr * dev(wq).to(pr.device)).sum().backward()
    np.testing.assert_allclose(m.grad.cpu().numpy(), fd, rtol=2e-6, atol=2e-8)


def test_product_state_input_gradients():
    """Non-basis product-state inputs go through the dense conjugation (cov_dense) and its adjoint kernel."""
    n, B = 4, 3
    rng = np.random.default_rng(13)
    layout, params = _random_stream(n, B, rng)
    amp = rng.standard_normal((n, 2)) + 1j * rng.standard_normal((n, 2))
    amp /= np.linalg.norm(amp, axis=1, keepdims=True)
    wires = [2, 0]
    wq = rng.standard_normal((B, 4))
    words, coeffs = ["ZIII", "IXXI", "IIYZ", "XYZX", "IIIX"], rng.standard_normal(5)
    wh = rng.standard_normal(B)

    def f_probs(pp):
        r = gates.chain(_oracle_ops(layout, pp), n)
        return float((readout.analytic_probability(r, amp, wires) * wq).sum())

    def f_expval(pp):
        r = gates.chain(_oracle_ops(layout, pp), n)
        return float((readout.expval_pauli_sum(cv.evolved_extended_covariance(amp, r), words, coeffs) * wh).sum())

    for f, kind in ((f_probs, "probs"), (f_expval, "expval")):
        fd = finite_difference(f, params)
        ptens = [dev(params[j], grad=True) for j in range(len(params))]
        d = mc.NIFDevice(n)
        from matchcake_b200.operations import ProductState
        d.execute_generator([ProductState(amp, wires=range(n))] + _device_ops(layout, ptens), reset=True)
        if kind == "probs":
            out = d.analytic_probability(wires)
            (out * dev(wq).to(out.device)).sum().backward()
        else:
            h = obs.Hamiltonian(list(coeffs), [obs.PauliWord({i: c for i, c in enumerate(w)}) for w in words])
            out = d.expval(h)
            (out * dev(wh).to(out.device)).sum().backward()
        got = np.stack([t.grad.cpu().numpy() for t in ptens])
        np.testing.assert_allclose(got, fd, rtol=2e-6, atol=2e-8)


def test_dense_sptm_circuit_gradients_through_device():
    """The reference's SPTM-circuit gradchecks (tests/test_devices/test_nif_device/test_gradients/
    test_apply_mths_gradients.py:156-260): a dense SPTM with requires_grad behind a rotation gate, probabilities and a
    Pauli-sum expectation value differentiated w.r.t. every matrix entry (covariance formulation, see the block test)."""
    n, B = 3, 2
    rng = np.random.default_rng(5)
    r0 = gates.random_sptm(n, seed=3, batch_size=B)
    p = rng.uniform(0, 6, (B, 2))
    bits = np.array([1, 0, 1])
    amp = cv.amplitudes_from_basis_state(bits)
    wq = rng.standard_normal((B, 4))
    states = readout.states_to_binary(np.arange(4), 2)
    words, coeffs = ["ZII", "IZZ", "XXI", "IYY", "XZX"], rng.standard_normal(5)
    wh = rng.standard_normal(B)

    def chain_of(m):
        return gates.chain([Op("SptmCompRxRx", (0, 1), params=p), Op("Sptm", (0, 1, 2), matrix=m)], n)

    def f_probs(m):
        pr = readout.probs_product_state(cv.evolve(cv.covariance_matrix(amp), chain_of(m)), states, np.array([0, 2])).T
        return float((pr / pr.sum(axis=-1, keepdims=True) * wq).sum())

    def f_expval(m):
        return float((readout.expval_pauli_sum(cv.evolved_extended_covariance(amp, chain_of(m)), words, coeffs) * wh).sum())

    for f, kind in ((f_probs, "probs"), (f_expval, "expval")):
        fd = finite_difference(f, r0)
        m = dev(r0, grad=True)
        d = mc.NIFDevice(n)
        d.execute_generator([BasisState(bits, wires=range(n)), SptmCompRxRx(dev(p), wires=[0, 1]),
                             SingleParticleTransitionMatrixOperation(m, wires=range(n))], reset=True)
        if kind == "probs":
            out = d.analytic_probability([0, 2])
            (out * dev(wq).to(out.device)).sum().backward()
        else:
            h = obs.Hamiltonian(list(coeffs), [obs.PauliWord({i: c for i, c in enumerate(w)}) for w in words])
            out = d.expval(h)
            (out * dev(wh).to(out.device)).sum().backward()
        np.testing.assert_allclose(m.grad.cpu().numpy(), fd, rtol=2e-6, atol=2e-8)


# ------------------------------------------------------------------ kernel alignment
def test_kernel_alignment_gradient_and_ascent():
    from matchcake_b200.ml.kernels import FermionicPQCKernel

    rng = np.random.default_rng(0)
    x = rng.random((24, 6))
    y = (x[:, 0] + x[:, 3] > 1.0).astype(int)
    k = FermionicPQCKernel(n_qubits=4, rotations="X,Z", r_dtype=torch.float64, c_dtype=torch.complex128)
    k.fit(x, y)
    c = k._create_kernel_centerer()
    yc = c @ k._create_y_kernel() @ c

    # gradient of the alignment w.r.t. bias_ vs finite differences of the (inference) forward path
    k.train()
    score = k.kernel_alignment(k(x, x), yc)
    score.backward()
    grad = k.bias_.grad.detach().cpu().numpy().copy()
    k.eval()
    b0 = k.bias_.detach().clone()

    def f(b):
        with torch.no_grad():
            k.bias_.copy_(torch.as_tensor(b, device="cuda"))
            k.float32_gram = False
            return float(k.kernel_alignment(k(x, x), yc))

    fd = finite_difference(f, b0.cpu().numpy())
    with torch.no_grad():
        k.bias_.copy_(b0)
    np.testing.assert_allclose(grad, fd, rtol=5e-6, atol=1e-8)

    # the optimisation loop improves the score and keeps the best parameters
    k2 = FermionicPQCKernel(n_qubits=4, rotations="X,Z", r_dtype=torch.float64, c_dtype=torch.complex128,
                            alignment=True, alignment_iterations=15, alignment_learning_rate=5e-2)
    k2.fit(x, y)
    hist = k2.opt_.history
    assert k2.opt_.fun == max(hist) and k2.opt_.fun > hist[0] + 1e-4
    assert not k2.training
    with torch.no_grad():
        k2.float32_gram = False
        final = float(k2.kernel_alignment(k2(x, x), yc))
    np.testing.assert_allclose(final, k2.opt_.fun, rtol=1e-9)


@pytest.mark.parametrize("n,m", [(6, 12), (40, 80), (96, 192)])
def test_cov_basis_backward_matches_torch_autograd(n, m):
    """Adjoint of cov = X^T Lambda_0 X: shared-memory kernel and (beyond its panel limit, (96, 192)) the global-memory
    variant against torch autograd over the same expression."""
    from matchcake_b200 import ops as mops

    torch.manual_seed(n)
    B, d = 3, 2 * n
    x = torch.randn(B, d, m, dtype=torch.float64, device="cuda")
    bits = torch.randint(0, 2, (n,), device="cuda")
    w = torch.randn(B, m, m, dtype=torch.float64, device="cuda")
    xa = x.clone().requires_grad_(True)
    (mops.cov_basis(xa, n, bits) * w).sum().backward()
    xb = x.clone().requires_grad_(True)
    sg = (bits.double() * 2 - 1)                                      # +1 for bit 1, -1 for bit 0
    lam0 = torch.zeros(d, d, dtype=torch.float64, device="cuda")
    ev = torch.arange(0, d, 2, device="cuda")
    lam0[ev, ev + 1] = sg
    lam0[ev + 1, ev] = -sg
    ref = xb.transpose(1, 2) @ lam0 @ xb
    np.testing.assert_allclose(mops.cov_basis(x, n, bits).cpu().numpy(), ref.detach().cpu().numpy(), rtol=1e-12, atol=1e-12)
    (ref * w).sum().backward()
    np.testing.assert_allclose(xa.grad.cpu().numpy(), xb.grad.cpu().numpy(), rtol=1e-11, atol=1e-11)


def _ry_matchgate(p):
    """M(R_y(theta), R_y(phi)) as a differentiable complex (..., 4, 4) tensor (comp_rotations.py:18-80)."""
    th, ph = p[..., 0], p[..., 1]
    z = torch.zeros_like(th)
    co, so, ci, si = torch.cos(th / 2), torch.sin(th / 2), torch.cos(ph / 2), torch.sin(ph / 2)
    rows = [torch.stack([co, z, z, -so], -1), torch.stack([z, ci, -si, z], -1), torch.stack([z, si, ci, z], -1),
            torch.stack([so, z, z, co], -1)]
    return torch.stack(rows, -2).to(torch.complex128)


@pytest.mark.parametrize("batched", [False, True])
def test_generic_matchgate_parameters_are_differentiable(batched):
    """Gradients flow through MatchgateOperation(U(p)) -- mcb200_matchgate_sptm and its adjoint kernel -- and agree with
    the closed-form rotation gate built from the same parameters (whose adjoint is checked against the reference's
    autograd above); products of matchgates on the same wires stay inside autograd."""
    import matchcake_b200 as mc
    from matchcake_b200 import observables as obs
    from matchcake_b200.operations import CompRxRx, CompRyRy, MatchgateOperation, fSWAP

    n = 4
    rng = np.random.default_rng(5)
    shape = (3, 2) if batched else (2,)
    p0 = torch.as_tensor(rng.uniform(0.2, 2.5, shape), device="cuda")
    q0 = torch.as_tensor(rng.uniform(0.2, 2.5, shape), device="cuda")
    proj = obs.Projector(np.array([0, 1, 1, 0]), wires=range(n))
    other = rng.uniform(0.2, 2.5, (2,))

    def run(generic):
        p = p0.clone().requires_grad_(True)
        q = q0.clone().requires_grad_(True)
        if generic:
            g1 = MatchgateOperation(_ry_matchgate(p), wires=[0, 1])
            g2 = MatchgateOperation(_ry_matchgate(q), wires=[2, 3]) @ MatchgateOperation(_ry_matchgate(p), wires=[2, 3])
            tail = [g1, g2]
        else:
            tail = [CompRyRy(p, wires=[0, 1]), CompRyRy(p, wires=[2, 3]), CompRyRy(q, wires=[2, 3])]
        stream = [CompRxRx(other, wires=[0, 1]), CompRxRx(other, wires=[2, 3]), fSWAP(wires=[1, 2])] + tail + \
                 [fSWAP(wires=[1, 2]), CompRxRx(other, wires=[1, 2])]
        dev = mc.NIFDevice(n, output_device="cuda")
        e = dev.execute_generator(stream, observable=proj, output_type="expval", reset=True)
        e.sum().backward()
        return e.detach().cpu().numpy(), p.grad.cpu().numpy(), q.grad.cpu().numpy()

    e_g, gp_g, gq_g = run(True)
    e_r, gp_r, gq_r = run(False)
    np.testing.assert_allclose(e_g, e_r, rtol=1e-11, atol=1e-13)
    assert np.abs(gp_r).max() > 1e-3 and np.abs(gq_r).max() > 1e-3
    np.testing.assert_allclose(gp_g, gp_r, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(gq_g, gq_r, rtol=1e-9, atol=1e-12)


def test_matchgate_sptm_operator_gradcheck():
    u = _ry_matchgate(torch.tensor([[0.4, 1.3], [2.0, 0.7]], dtype=torch.float64, device="cuda"))
    ur = torch.view_as_real(u).contiguous().requires_grad_(True)
    assert torch.autograd.gradcheck(lambda t: torch.ops.mcb200.matchgate_sptm(t)[0], (ur,), atol=1e-7, rtol=1e-6)
