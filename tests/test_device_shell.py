"""PennyLane-facing device entry points on tape stubs (PennyLane is absent from the image), star-state search, shot-based
read-outs and LinearNIFKernel -- reference: devices/nif_device.py:562-578,864-883,949-989,1057-1115,
devices/star_state_finding_strategies/, ml/kernels/linear_nif_kernel.py."""
import dataclasses
import itertools

import numpy as np
import pytest
import torch

import matchcake_b200 as mc
from matchcake_b200 import observables as obs
from matchcake_b200.ml.kernels import LinearNIFKernel
from matchcake_b200.operations import (BasisState, CompRxRx, CompRyRy, CompRzRz, MatchgateOperation, SptmCompRxRx,
                                       SptmCompRyRy, SptmFSwap, fSWAP)
from matchcake_b200.pennylane_shell import ExecutionConfig, HostPipeline, HostTape
from matchcake_b200.star_state import (FromSamplingStrategy, GreedyStrategy, StarStateFindingStrategy,
                                       get_star_state_finding_strategy)
from oracle import gates, kernels, readout
from oracle.gates import Op

RTOL, ATOL = 1e-10, 1e-13


# measurement records recognised by class name, like pennylane.measurements.*
class ExpectationMP:
    def __init__(self, obs):
        self.obs, self.wires = obs, getattr(obs, "wires", ())


class ProbabilityMP:
    def __init__(self, wires=None):
        self.obs, self.wires = None, tuple(wires or ())


class SampleMP:
    def __init__(self):
        self.obs, self.wires = None, ()


class CountsMP:
    obs, wires = None, ()


# ------------------------------------------------------------------------------------------------ host logic (CPU)
def test_setup_execution_config_defaults_to_backprop():
    dev = mc.NIFDevice(3)
    assert dev.setup_execution_config().gradient_method == "backprop"
    assert dev.setup_execution_config(ExecutionConfig(gradient_method="best")).gradient_method == "backprop"
    cfg = ExecutionConfig(gradient_method="parameter-shift", interface="torch")
    out = dev.setup_execution_config(cfg)
    assert out.gradient_method == "parameter-shift" and out.interface == "torch"


def test_supports_derivatives_rules():
    dev = mc.NIFDevice(3)
    assert dev.supports_derivatives() is True
    assert dev.supports_derivatives(ExecutionConfig(gradient_method="backprop")) is True
    assert dev.supports_derivatives(ExecutionConfig(gradient_method="best")) is True
    assert dev.supports_derivatives(ExecutionConfig(gradient_method="adjoint")) is False
    analytic = HostTape([], [], shots=None)
    sampled = HostTape([], [], shots=100)
    assert dev.supports_derivatives(ExecutionConfig(gradient_method="backprop"), analytic) is True
    assert dev.supports_derivatives(ExecutionConfig(gradient_method="backprop"), sampled) is False


def test_stopping_condition_and_host_pipeline_decomposition():
    dev = mc.NIFDevice(4)
    assert dev._stopping_condition(CompRxRx(np.zeros(2), wires=[0, 1]))
    assert dev._stopping_condition(SptmFSwap(wires=[1, 2]))
    assert dev._stopping_condition(BasisState([0, 0, 0, 0], wires=range(4)))
    assert "SptmCompRzRz" in dev._supported_ops and "CompZX" in dev._supported_ops

    class Unsupported:
        name = "Unsupported"
        wires = (0, 1)

        def matrix(self):
            raise RuntimeError("no matrix")

    class Composite(Unsupported):
        name = "Composite"

        def decomposition(self):
            return [CompRyRy(np.array([0.1, 0.2]), wires=[0, 1]), Nested()]

    class Nested(Unsupported):
        name = "Nested"

        def decomposition(self):
            return [fSWAP(wires=[1, 2])]

    assert not dev._stopping_condition(Unsupported())
    pipe = dev.preprocess_transforms()
    assert isinstance(pipe, HostPipeline) and len(pipe) == 2
    tape = HostTape([Composite(), CompRxRx(np.zeros(2), wires=[2, 3])], [ProbabilityMP(wires=[0])], shots=None)
    (new,), post = pipe(tape)
    assert [type(o).__name__ for o in new.operations] == ["CompRyRy", "CompZX", "CompRxRx"]
    assert post(("r",)) == "r"
    # strict=False: an operation without decomposition is left for the device to reject
    (left,), _ = pipe(HostTape([Unsupported()], []))
    assert type(left.operations[0]).__name__ == "Unsupported"


def test_star_state_strategy_lookup():
    assert isinstance(get_star_state_finding_strategy("greedy"), GreedyStrategy)
    assert isinstance(get_star_state_finding_strategy(" FromSampling "), FromSamplingStrategy)
    g = GreedyStrategy()
    assert get_star_state_finding_strategy(g) is g
    with pytest.raises(ValueError):
        get_star_state_finding_strategy("a*")
    assert issubclass(GreedyStrategy, StarStateFindingStrategy)
    assert isinstance(mc.NIFDevice(3).star_state_finding_strategy, FromSamplingStrategy)
    assert isinstance(mc.NIFDevice(3, star_state_finding_strategy="Greedy").star_state_finding_strategy, GreedyStrategy)


class _TableDevice:
    """Host stand-in with an explicit probability table p[b, state]: what GreedyStrategy sees of a device."""

    def __init__(self, table):
        self.table = np.asarray(table, dtype=float)  # (B, 2^n)
        self.num_wires = int(np.log2(self.table.shape[1]))
        self.wires = tuple(range(self.num_wires))
        self.calls = []

    states_to_binary = staticmethod(mc.NIFDevice.states_to_binary)

    def prob(self, states, wires):
        states = np.atleast_2d(states)
        self.calls.append(states.shape)
        n, k = self.num_wires, states.shape[1]
        full = self.states_to_binary(np.arange(2 ** n), n)
        out = np.stack([self.table[:, (full[:, :k] == s).all(1)].sum(1) for s in states])   # (Q, B) marginals
        return out if self.table.shape[0] > 1 else out[:, 0]


@pytest.mark.parametrize("n,B", [(4, 1), (5, 1), (6, 3), (7, 4)])
def test_greedy_strategy_follows_the_reference_search(n, B):
    rng = np.random.default_rng(n * 10 + B)
    table = rng.random((B, 2 ** n)) ** 4
    table /= table.sum(1, keepdims=True)
    dev = _TableDevice(table)
    states, probs = GreedyStrategy()(dev, dev.prob)
    # restatement of greedy_strategy.py:23-52 with duplicated prefixes (unique=False), on the same table
    blocks = dev.states_to_binary(np.arange(4), 2)
    p = dev.prob(blocks, None)
    ref_states, ref_probs = blocks[np.argmax(p, axis=0)], np.max(p, axis=0)
    left = n % 2
    for kj in [2] * len(range(2, n - left, 2)) + ([left] if left else []):
        add = dev.states_to_binary(np.arange(2 ** kj), kj)
        pre = ref_states.reshape(-1, ref_states.shape[-1])
        ext = np.concatenate([np.tile(pre, (len(add), 1)), np.repeat(add, len(pre), axis=0)], axis=1)  # block-major
        p = dev.prob(ext, None)
        ref_states, ref_probs = ext[np.argmax(p, axis=0)], np.max(p, axis=0)
    np.testing.assert_array_equal(states, ref_states)
    np.testing.assert_array_equal(probs, ref_probs)
    assert states.shape == ((n,) if B == 1 else (B, n))


def test_from_sampling_strategy_counts_the_mode():
    samples = np.array([[[0, 1], [1, 1]], [[0, 1], [0, 0]], [[1, 0], [1, 1]], [[0, 1], [1, 1]]])  # (shots=4, B=2, n=2)

    class Dev:
        pass

    d = Dev()
    d.samples = samples
    s, p = FromSamplingStrategy()(d, None)
    np.testing.assert_array_equal(s, [[0, 1], [1, 1]])
    np.testing.assert_allclose(p, [0.75, 0.75])
    s1, p1 = FromSamplingStrategy()(d, None, samples=samples[:, 0])
    np.testing.assert_array_equal(s1, [0, 1])
    assert p1 == 0.75


def test_sampled_read_outs_from_recorded_samples():
    """process_samples semantics on fixed samples: no GPU involved once the samples exist."""
    dev = mc.NIFDevice(3, shots=8)
    dev._samples = np.array([[0, 0, 1], [0, 0, 1], [1, 0, 1], [1, 1, 0], [0, 0, 1], [0, 1, 1], [1, 0, 1], [0, 0, 0]])
    p = dev.probability(wires=[0, 2]).numpy()
    np.testing.assert_allclose(p, [1 / 8, 4 / 8, 1 / 8, 2 / 8])
    np.testing.assert_allclose(float(dev.expval(obs.Projector([0, 0, 1], wires=[0, 1, 2]))), 3 / 8)
    zz = obs.Z(0) @ obs.Z(2)
    np.testing.assert_allclose(float(dev.expval(0.5 * zz + obs.PauliSum([2.0], [obs.Z(1)]))),
                               0.5 * np.mean([-1, -1, 1, -1, -1, -1, 1, 1]) + 2.0 * np.mean([1, 1, 1, -1, 1, -1, 1, 1]))
    np.testing.assert_allclose(dev.expval(obs.Z(0), bin_size=4).numpy(), [0.0, 0.5])
    np.testing.assert_allclose(float(dev.expval(obs.Z(0), shot_range=(0, 2))), 1.0)
    with pytest.raises(mc.DeviceError):
        dev.expval(obs.X(0))


@pytest.mark.parametrize("n_qubits,act,bias", [(3, "Identity", True), (4, "Tanh", False)])
def test_linear_nif_kernel_host_surface(n_qubits, act, bias):
    """reference tests/test_ml/test_kernels/test_linear_nif_kernel.py: dtype propagation and the three setters."""
    k = LinearNIFKernel(n_qubits=n_qubits, bias=bias, encoder_activation=act, r_dtype=torch.float64,
                        c_dtype=torch.complex128)
    assert k.R_DTYPE == torch.float64 and k.C_DTYPE == torch.complex128
    assert k.q_device.R_DTYPE == torch.float64 and k.q_device.C_DTYPE == torch.complex128
    d = 2 * n_qubits
    assert k.encoder[1].out_features == d * (d - 1) // 2 == k.encoder_out_indices[0].size
    k.n_qubits = n_qubits + 1
    assert k.n_qubits == n_qubits + 1 and k.q_device.num_wires == n_qubits + 1
    assert k.encoder[1].out_features == k.encoder_out_indices[0].size == (d + 2) * (d + 1) // 2
    k.bias = not bias
    assert k.bias == (not bias)
    k.encoder_activation = "ReLU"
    assert k.encoder_activation == "ReLU" and isinstance(k.encoder[-1], torch.nn.ReLU)
    params = k.get_params()
    assert params["n_qubits"] == n_qubits + 1 and params["encoder_activation"] == "ReLU"


# ------------------------------------------------------------------------------------------------ on the GPU
def _readme_stream(p, n):
    ops = []
    for q in range(0, n - 1, 2):
        ops += [CompRxRx(p, wires=[q, q + 1]), CompRyRy(p, wires=[q, q + 1]), CompRzRz(p, wires=[q, q + 1])]
    for q in range(1, n - 1, 2):
        ops.append(fSWAP(wires=[q, q + 1]))
    return ops


def _readme_oracle(p, n):
    ops = []
    for q in range(0, n - 1, 2):
        ops += [Op(k, (q, q + 1), params=p) for k in ("CompRxRx", "CompRyRy", "CompRzRz")]
    for q in range(1, n - 1, 2):
        ops.append(Op("fSWAP", (q, q + 1)))
    return gates.chain(ops, n)


@pytest.mark.gpu
def test_execute_on_tape_stubs_matches_oracle():
    n, B = 6, 7
    rng = np.random.default_rng(0)
    p = rng.uniform(0, 2 * np.pi, (B, 2))
    r = _readme_oracle(p, n)
    zeros = np.zeros(n, dtype=int)
    tgt = np.array([0, 1, 1, 0, 0, 0])
    proj = obs.Projector(tgt, wires=range(n))
    dev = mc.NIFDevice(n)
    tape = HostTape(_readme_stream(p, n), [ExpectationMP(proj), ProbabilityMP(wires=[1, 4])])
    e, pr = dev.execute(tape)
    np.testing.assert_allclose(e.numpy(), readout.probs_lookup_table(r, zeros, tgt, np.arange(n)), rtol=RTOL, atol=ATOL)
    from oracle import covariance as cv

    np.testing.assert_allclose(pr.numpy(), readout.analytic_probability(r, cv.amplitudes_from_basis_state(zeros), [1, 4]),
                               rtol=RTOL, atol=ATOL)
    # a batch of circuits -> a tuple of results; a single measurement -> the bare result
    res = dev.execute([HostTape(_readme_stream(p[:2], n), [ExpectationMP(proj)]),
                       HostTape(_readme_stream(p[2:5], n), [ExpectationMP(proj)])], ExecutionConfig())
    assert isinstance(res, tuple) and len(res) == 2 and res[0].shape == (2,) and res[1].shape == (3,)
    np.testing.assert_allclose(torch.cat(res).numpy(), e.numpy()[:5], rtol=1e-12)
    with pytest.raises(NotImplementedError):
        dev.execute(HostTape(_readme_stream(p, n), [CountsMP()]))
    # the preprocessing pipeline feeds execute
    tapes, post = dev.preprocess_transforms()(tape)
    np.testing.assert_allclose(post(dev.execute(tapes))[0].numpy(), e.numpy(), rtol=0, atol=0)


@pytest.mark.gpu
def test_execute_with_shots_returns_samples_and_estimates():
    n = 4
    p = np.array([0.7, 1.9])
    dev = mc.NIFDevice(n, seed=3)
    proj = obs.Projector(np.zeros(n, dtype=int), wires=range(n))
    tape = HostTape(_readme_stream(p, n), [SampleMP(), ExpectationMP(proj), ProbabilityMP(wires=[0, 1])], shots=20000)
    s, e, pr = dev.execute(tape)
    assert s.shape == (20000, n) and set(np.unique(s)) <= {0, 1}
    exact = mc.NIFDevice(n)
    exact.execute_generator(_readme_stream(p, n), reset=True)
    p_exact = float(exact.exact_expval(proj))
    assert abs(float(e) - p_exact) < 5 * np.sqrt(p_exact * (1 - p_exact) / 20000) + 1e-3
    np.testing.assert_allclose(pr.numpy(), exact.analytic_probability([0, 1]).numpy(), atol=0.02)
    assert abs(float(pr.sum()) - 1.0) < 1e-12
    # execute_generator with a per-call shots override takes the sampled branch as well (nif_device.py:657-668)
    e2 = exact.execute_generator(_readme_stream(p, n), observable=proj, output_type="expval", reset=True, shots=5000)
    assert abs(float(e2) - p_exact) < 0.05 and float(e2) * 5000 == round(float(e2) * 5000)


@pytest.mark.gpu
@pytest.mark.parametrize("n,B", [(4, None), (5, None), (6, 5)])
def test_greedy_star_state_on_the_device(n, B):
    rng = np.random.default_rng(n)
    p = rng.uniform(0, 2 * np.pi, (2,) if B is None else (B, 2))
    dev = mc.NIFDevice(n, star_state_finding_strategy="Greedy")
    states, probs = dev.execute_generator(_readme_stream(p, n), output_type="star_state", reset=True)
    r = _readme_oracle(p, n)
    zeros = np.zeros(n, dtype=int)
    all_states = readout.states_to_binary(np.arange(2 ** n), n)
    table = readout.probs_lookup_table(r, zeros, all_states, np.arange(n))   # (2^n,) or (2^n, B)
    table = table.reshape(2 ** n, -1).T
    host = _TableDevice(table)
    ref_states, ref_probs = GreedyStrategy()(host, host.prob)
    np.testing.assert_array_equal(states, ref_states)
    np.testing.assert_allclose(probs, ref_probs, rtol=1e-9, atol=1e-12)
    assert dev.star_state is states and dev.star_probability is probs
    # the probability attached to the star state is its exact probability
    got = np.atleast_2d(states)
    idx = (got << np.arange(n - 1, -1, -1)).sum(-1)
    np.testing.assert_allclose(np.atleast_1d(probs), table[np.arange(len(idx)), idx], rtol=1e-9)


@pytest.mark.gpu
def test_from_sampling_star_state_on_the_device():
    n = 4
    dev = mc.NIFDevice(n, shots=4000, seed=1)
    state, prob = dev.execute_generator(_readme_stream(np.array([0.3, 0.2]), n), output_type="star_state", reset=True)
    np.testing.assert_array_equal(state, np.zeros(n, dtype=int))   # small angles: |0000> dominates
    assert 0.5 < float(prob) <= 1.0


@pytest.mark.gpu
def test_three_batched_matchgates_in_one_segment_keep_their_own_parameters():
    """parameter columns of matrices derived on the fly must not be shared through a recycled id()."""
    rng = np.random.default_rng(11)
    n, B = 8, 4
    mats, stream, oops = [], [], []
    for j, w in enumerate([0, 2, 4, 6, 1, 3, 5]):
        u = gates.matchgate_comp_rotation(rng.uniform(0, 6, (B, 2)), "XX") @ \
            gates.matchgate_comp_rotation(rng.uniform(0, 6, (B, 2)), "YY")
        mats.append(u)
        stream.append(MatchgateOperation(u, wires=[w, w + 1]))
        oops.append(Op("Matchgate", (w, w + 1), matrix=u))
    dev = mc.NIFDevice(n)
    dev.execute_generator(stream, reset=True)
    np.testing.assert_allclose(dev.global_sptm.matrix().cpu().numpy(), gates.chain(oops, n), rtol=RTOL, atol=ATOL)


@pytest.mark.gpu
@pytest.mark.parametrize("n_qubits,act,bias", [(3, "Identity", True), (4, "Tanh", False), (8, "Tanh", True)])
def test_linear_nif_kernel_gram_matches_oracle(n_qubits, act, bias):
    rng = np.random.default_rng(n_qubits)
    x = rng.random((10, 8)) * 0.5
    torch.manual_seed(n_qubits)
    k = LinearNIFKernel(n_qubits=n_qubits, bias=bias, encoder_activation=act, r_dtype=torch.float64,
                        c_dtype=torch.complex128)
    k.float32_gram = False
    k.fit(x)
    got = k(x).cpu().numpy()
    lin = k.encoder[1]
    o = kernels.LinearNIFKernelOracle(n_qubits, lin.weight.detach().cpu().numpy(),
                                      None if lin.bias is None else lin.bias.detach().cpu().numpy(), act).fit(x)
    ref = o.gram(x, float32_store=False)
    np.testing.assert_allclose(got, ref, rtol=1e-9, atol=0)
    np.testing.assert_array_equal(got, got.T)
    # the SPTMs are the matrix exponentials of the encoder's generator
    np.testing.assert_allclose(k._x_to_sptm(x).cpu().numpy(), o.x_to_sptm(x), rtol=0, atol=1e-13)
    # reference test_swap_test: identical rows -> all ones
    same = np.stack([np.linspace(0.0, 1.0, num=8) for _ in range(6)])
    k2 = LinearNIFKernel(n_qubits=n_qubits, bias=bias, encoder_activation=act).fit(same)
    km = k2(same)
    torch.testing.assert_close(km, torch.ones_like(km))


@pytest.mark.gpu
def test_linear_nif_kernel_encoder_is_trainable_through_the_adjoint_kernels():
    rng = np.random.default_rng(0)
    x = rng.random((6, 5))
    torch.manual_seed(0)
    k = LinearNIFKernel(n_qubits=3, r_dtype=torch.float64, c_dtype=torch.complex128)
    k.float32_gram = False
    k.fit(x)
    k.train()
    gram = k(x)
    assert gram.requires_grad
    gram.sum().backward()
    w = k.encoder[1].weight
    assert w.grad is not None and float(w.grad.abs().max()) > 0
    # finite difference of one weight through the oracle-free forward
    with torch.no_grad():
        k.eval()
        eps = 1e-5
        w[2, 1] += eps
        up = float(k(x).double().sum())
        w[2, 1] -= 2 * eps
        down = float(k(x).double().sum())
        w[2, 1] += eps
    assert abs((up - down) / (2 * eps) - float(w.grad[2, 1])) < 1e-6 * max(1.0, abs(float(w.grad[2, 1])))


@pytest.mark.gpu
@pytest.mark.parametrize("host", [True, False])
def test_execute_layered_equals_the_op_stream(host):
    """One (B, P) parameter tensor + layout == the same circuit given as operation objects, for streamed host parameters
    (several chunks, ragged last chunk) and for device-resident ones; probabilities and projector read-out."""
    n, depth, B, k = 10, 6, 37, 3
    rng = np.random.default_rng(5)
    classes = (SptmCompRxRx, SptmCompRyRy, CompRzRz)
    layout = []
    for l in range(depth):
        for q in range(l % 2, n - 1, 2):
            layout.append((classes[l % 3], q))
        if l == 2:
            layout.append((SptmFSwap, (7, 2)))
            layout.append((fSWAP, (4, 5)))
    n_rot = sum(1 for cls, _ in layout if cls not in (SptmFSwap, fSWAP))
    params = rng.uniform(0, 2 * np.pi, (B, 2 * n_rot))
    dev = mc.NIFDevice(n)
    p_in = torch.as_tensor(params).pin_memory() if host else torch.as_tensor(params).cuda()
    got = dev.execute_layered(layout, p_in, probs_wires=[0, 4, 9], chunk_rows=8 if host else None)
    stream, j = [], 0
    for cls, w in layout:
        if cls in (SptmFSwap, fSWAP):
            stream.append(cls(wires=list(w)))
        else:
            stream.append(cls(params[:, 2 * j:2 * j + 2], wires=[w, w + 1]))
            j += 1
    ref_dev = mc.NIFDevice(n)
    ref_dev.execute_generator(stream, reset=True)
    ref = ref_dev.analytic_probability([0, 4, 9])
    assert got.shape == (B, 8) and got.device.type == "cpu"
    np.testing.assert_array_equal(got.numpy(), ref.numpy())
    tgt = np.array([1, 0, 1, 1])
    proj = obs.Projector(tgt, wires=[2, 3, 5, 8])
    got_e = dev.execute_layered(layout, p_in, observable=proj, chunk_rows=16 if host else None)
    np.testing.assert_array_equal(got_e.numpy(), ref_dev.expval(proj).numpy())
    with pytest.raises(ValueError):
        dev.execute_layered(layout, p_in[:, :-2], probs_wires=[0])
    with pytest.raises(ValueError):
        dev.execute_layered(layout, p_in)
    with pytest.raises(mc.DeviceError):
        dev.execute_layered([(MatchgateOperation, 0)], p_in, probs_wires=[0])


@pytest.mark.gpu
def test_streamed_host_parameters_equal_device_resident_ones():
    """Pinned host parameters of a large batch go through the fused projector kernel in chunks (copy / compute overlap):
    same values as the device-resident batch, bit for bit; later read-outs of the same program still work."""
    n, B = 8, 20001   # odd batch: ragged last chunk
    rng = np.random.default_rng(3)
    p = rng.uniform(0, 2 * np.pi, (B, 2))
    proj = obs.Projector(np.zeros(n, dtype=int), wires=range(n))
    dev_s = mc.NIFDevice(n)
    dev_s.STREAM_MIN_BYTES, dev_s.STREAM_CHUNK_BYTES = 1 << 16, 96 * 1024   # force the streamed path: 4 chunks of 6144 rows
    host = torch.as_tensor(p).pin_memory()
    got = dev_s.execute_generator(_readme_stream(host, n), observable=proj, output_type="expval", reset=True)
    dev_d = mc.NIFDevice(n)
    ref = dev_d.execute_generator(_readme_stream(torch.as_tensor(p).cuda(), n), observable=proj, output_type="expval",
                                  reset=True)
    assert got.device.type == "cpu" and got.shape == (B,)
    np.testing.assert_array_equal(got.numpy(), ref.numpy())
    sel = rng.choice(B, 64, replace=False)
    zeros = np.zeros(n, dtype=int)
    np.testing.assert_allclose(got.numpy()[sel], readout.probs_lookup_table(_readme_oracle(p[sel], n), zeros, zeros, np.arange(n)),
                               rtol=RTOL, atol=ATOL)
    # the same (still host-resident) program serves other read-outs
    pr = dev_s.analytic_probability([0, 1])
    np.testing.assert_array_equal(pr.numpy(), dev_d.analytic_probability([0, 1]).numpy())
