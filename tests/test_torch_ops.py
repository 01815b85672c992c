"""``torch.ops.mcb200.*``: every compute entry point of the C ABI is a registered torch custom operator with a fake
(shape) implementation and, where the path is differentiable, a registered autograd formula."""
import numpy as np
import pytest
import torch

import matchcake_b200  # noqa: F401  (registers the operators)
from matchcake_b200 import _lib
from matchcake_b200 import ops as mops
from matchcake_b200 import torch_ops as T

f64 = torch.float64

# C entry point -> operator(s) that launch it.  Library management / option / probe entry points have no tensor
# arguments and stay plain ctypes calls.
ENTRY_TO_OP = {
    "mcb200_circuit_columns": "circuit_columns", "mcb200_sptm_matmul": "sptm_matmul",
    "mcb200_matchgate_sptm": "matchgate_sptm", "mcb200_product_state_cov": "product_state_cov",
    "mcb200_cov_basis": "cov_basis", "mcb200_cov_dense": "cov_dense", "mcb200_pfaffian": "pfaffian",
    "mcb200_probs": "probs", "mcb200_probs_all": "probs_all", "mcb200_sector_features": "sector_features",
    "mcb200_rowdot": "rowdot", "mcb200_pauli_sum_small": "pauli_sum_small",
    "mcb200_circuit_expval_projector": "circuit_expval_projector", "mcb200_circuit_cov_basis": "circuit_cov_basis", "mcb200_gram": "gram_out",
    "mcb200_circuit_pair_blocks": "circuit_pair_blocks", "mcb200_gram_blocks4": "gram_blocks4_out",
    "mcb200_gram_symmetrize": "gram_symmetrize_", "mcb200_sample": "sample",
    "mcb200_matchgate_sptm_backward": "matchgate_sptm_backward", "mcb200_pfaffian_backward": "pfaffian_backward", "mcb200_probs_backward": "probs_backward",
    "mcb200_sector_features_backward": "sector_features_backward", "mcb200_cov_basis_backward": "cov_basis_backward",
    "mcb200_cov_dense_backward": "cov_dense_backward", "mcb200_circuit_columns_backward": "circuit_columns_backward",
}
NOT_OPS = {"mcb200_circuit_validate", "mcb200_version", "mcb200_last_error", "mcb200_device_count", "mcb200_set_abs_floor", "mcb200_get_abs_floor",
           "mcb200_set_gram_pivot_growth", "mcb200_get_gram_pivot_growth",
           "mcb200_set_expval_route", "mcb200_launch_count", "mcb200_reset_launch_count", "mcb200_probe_fp64_tflops"}


def test_every_compute_entry_point_is_a_registered_operator():
    for sym in _lib.SIGNATURES:
        if sym in NOT_OPS or sym.endswith("_workspace_bytes"):
            continue
        assert sym in ENTRY_TO_OP, f"{sym} has no torch operator"
        name = ENTRY_TO_OP[sym]
        assert name in T.OPS
        op = getattr(torch.ops.mcb200, name)
        assert op.default._schema.name == f"mcb200::{name}"
    assert set(ENTRY_TO_OP.values()) | {"cov_basis_out"} == set(T.OPS)


def _meta(*shape, dtype=f64):
    return torch.empty(shape, dtype=dtype, device="meta")


def test_fake_implementations_propagate_shapes_without_a_gpu():
    """Meta tensors dispatch to the registered fake kernels: no launch, no library call."""
    o = torch.ops.mcb200
    circ = mops.CompiledCircuit(4, [mops.GateSpec(mops.GATE_RXRX, 0, 1, 0), mops.GateSpec(mops.GATE_FSWAP, 1, 2)], 2,
                                device="cpu")
    p = _meta(7, 2)
    assert o.circuit_columns(p, circ.handle, None).shape == (7, 8, 8)
    assert o.circuit_columns(p, circ.handle, _meta(3, dtype=torch.int32)).shape == (7, 8, 3)
    assert o.circuit_columns_backward(p, circ.handle, _meta(7, 8, 3), _meta(7, 8, 3)).shape == (7, 2)
    assert o.circuit_expval_projector(p, circ.handle, None, None, 0.0).shape == (7,)
    assert o.circuit_cov_basis(p, circ.handle, _meta(4, dtype=torch.int32), None).shape == (7, 4, 4)
    assert o.circuit_pair_blocks(p, circ.handle, _meta(2, dtype=torch.int32), None).shape == (7, 2, 6)
    assert o.sptm_matmul(_meta(5, 6, 6), _meta(6, 6), False, True).shape == (5, 6, 6)
    a, b = o.matchgate_sptm(_meta(9, 4, 4, 2))
    assert a.shape == (9, 4, 4) and b.shape == (1,)
    assert o.product_state_cov(_meta(3, 5, 2, 2), True).shape == (3, 11, 11)
    assert o.cov_basis(_meta(7, 8, 4), 4, None).shape == (7, 4, 4)
    assert o.cov_basis_backward(_meta(7, 8, 4), 4, None, _meta(7, 4, 4)).shape == (7, 8, 4)
    assert o.cov_dense(_meta(7, 8, 8), _meta(9, 9), _meta(4, dtype=torch.int32)).shape == (7, 4, 4)
    assert o.cov_dense_backward(_meta(8, 8), _meta(7, 9, 9), None, _meta(7, 9, 9)).shape == (1, 8, 8)
    assert o.pfaffian(_meta(11, 6, 6), False, 1.0, 0.0).shape == (11,)
    assert o.pfaffian_backward(_meta(11, 6, 6), _meta(11), _meta(11)).shape == (11, 6, 6)
    assert o.probs(_meta(7, 6, 6), _meta(5, 3, dtype=torch.int8), False, 0.0).shape == (7, 5)
    assert o.probs_all(_meta(7, 6, 6), True, 0.0).shape == (7, 8)
    bits, prob = o.sample(_meta(7, 6, 6), 10, -3, None, True)
    assert bits.shape == (10, 7, 3) and bits.dtype == torch.int8 and prob.shape == (10, 7)
    assert o.sector_features(_meta(7, 9, 9), _meta(12, 4, dtype=torch.int32)).shape == (7, 12)
    assert o.rowdot(_meta(7, 12), _meta(12), None).shape == (7,)
    assert o.pauli_sum_small(_meta(7, 9, 9), _meta(4, dtype=torch.int32), _meta(8, dtype=torch.int32), _meta(3),
                             None).shape == (7,)
    k = _meta(5, 6)
    assert o.gram_out(k, _meta(5, 4, 4), _meta(6, 4, 4), 0.25, 0, 0, 5, False, 0.0) is None
    assert o.gram_symmetrize_(k) is None


def test_circuit_handles_are_weak():
    circ = mops.CompiledCircuit(2, [mops.GateSpec(mops.GATE_RXRX, 0, 1, 0)], 2, device="cpu")
    h = circ.handle
    assert T.circuit(h) is circ
    del circ
    import gc

    gc.collect()
    with pytest.raises(_lib.Mcb200Error):
        T.circuit(h)


def test_seed_round_trip():
    for s in (0, 1, 2 ** 63 - 1, 2 ** 63, 2 ** 64 - 1, 0x9E3779B97F4A7C15):
        assert T.signed_seed(s) & 0xFFFFFFFFFFFFFFFF == s
        assert -(2 ** 63) <= T.signed_seed(s) < 2 ** 63


def test_cpu_tensors_have_no_kernel():
    """Only a CUDA kernel is registered: the dispatcher itself refuses CPU tensors (no fallback)."""
    with pytest.raises((NotImplementedError, RuntimeError)):
        torch.ops.mcb200.pfaffian(torch.zeros((1, 2, 2), dtype=f64), False, 1.0, 0.0)


# ------------------------------------------------------------------------------------------------ on the GPU
@pytest.mark.gpu
def test_opcheck_schema_fake_and_autograd_registration():
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    a = torch.randn(5, 8, 8, dtype=f64, device=dev, generator=g)
    a = (a - a.transpose(1, 2)).requires_grad_(True)
    tests = ("test_schema", "test_faketensor", "test_autograd_registration")
    torch.library.opcheck(torch.ops.mcb200.pfaffian.default, (a, True, 1.0, 0.0), test_utils=tests)
    x = torch.randn(4, 8, 6, dtype=f64, device=dev, generator=g).requires_grad_(True)
    torch.library.opcheck(torch.ops.mcb200.cov_basis.default, (x, 4, None), test_utils=tests)
    m1 = torch.randn(3, 8, 8, dtype=f64, device=dev, generator=g).requires_grad_(True)
    m2 = torch.randn(8, 8, dtype=f64, device=dev, generator=g).requires_grad_(True)
    torch.library.opcheck(torch.ops.mcb200.sptm_matmul.default, (m1, m2, False, True), test_utils=tests)
    circ = mops.CompiledCircuit(4, [mops.GateSpec(mops.GATE_RXRX, 0, 1, 0), mops.GateSpec(mops.GATE_RZRZ, 2, 3, 2),
                                    mops.GateSpec(mops.GATE_FSWAP, 1, 2)], 4)
    p = torch.rand(6, 4, dtype=f64, device=dev, generator=g).requires_grad_(True)
    torch.library.opcheck(torch.ops.mcb200.circuit_columns.default, (p, circ.handle, None), test_utils=tests)
    k = torch.zeros(5, 5, dtype=f64, device=dev)
    torch.library.opcheck(torch.ops.mcb200.gram_out.default,
                          (k, a.detach(), a.detach(), 0.0625, mops.GRAM_REFERENCE, 0, 5, False, 0.0),
                          test_utils=("test_schema", "test_faketensor"))


@pytest.mark.gpu
def test_operator_gradients_match_finite_differences():
    """gradcheck straight on the registered operators (the adjoint kernels behind register_autograd)."""
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(1)
    up = torch.randn(3, 6, 6, dtype=f64, device=dev, generator=g).requires_grad_(True)

    def pf(u):
        a = torch.triu(u, 1)
        return torch.ops.mcb200.pfaffian(a - a.transpose(1, 2), True, 1.0, 0.0)

    assert torch.autograd.gradcheck(pf, (up,), atol=1e-6, rtol=1e-5)
    m1 = torch.randn(2, 4, 4, dtype=f64, device=dev, generator=g).requires_grad_(True)
    m2 = torch.randn(4, 4, dtype=f64, device=dev, generator=g).requires_grad_(True)
    assert torch.autograd.gradcheck(lambda p, q: torch.ops.mcb200.sptm_matmul(p, q, True, False), (m1, m2),
                                    atol=1e-6, rtol=1e-5)
    f = torch.randn(4, 5, dtype=f64, device=dev, generator=g).requires_grad_(True)
    c = torch.randn(5, dtype=f64, device=dev, generator=g).requires_grad_(True)
    acc = torch.randn(4, dtype=f64, device=dev, generator=g).requires_grad_(True)
    assert torch.autograd.gradcheck(lambda a_, b_, c_: torch.ops.mcb200.rowdot(a_, b_, c_), (f, c, acc), atol=1e-6,
                                    rtol=1e-5)


@pytest.mark.gpu
def test_float32_parameters_requiring_grad_get_correct_gradients():
    """The wrapper casts to float64 differentiably; the adjoint kernel always sees the float64 contiguous copy."""
    dev = "cuda"
    circ = mops.CompiledCircuit(4, [mops.GateSpec(mops.GATE_RXRX, 0, 1, 0), mops.GateSpec(mops.GATE_RYRY, 2, 3, 2),
                                    mops.GateSpec(mops.GATE_FSWAP, 1, 2), mops.GateSpec(mops.GATE_RZRZ, 0, 1, 0)], 4)
    base = torch.rand(5, 4, dtype=f64, device=dev, generator=torch.Generator(device=dev).manual_seed(2))
    p64 = base.clone().requires_grad_(True)
    p32 = base.to(torch.float32).requires_grad_(True)
    w = torch.randn(5, 8, 8, dtype=f64, device=dev, generator=torch.Generator(device=dev).manual_seed(3))
    (mops.circuit_columns(circ, p64) * w).sum().backward()
    (mops.circuit_columns(circ, p32) * w).sum().backward()
    assert p32.grad.dtype == torch.float32
    np.testing.assert_allclose(p32.grad.double().cpu().numpy(), p64.grad.cpu().numpy(), rtol=1e-5, atol=1e-5)
    # non-contiguous float64 view
    wide = torch.zeros(5, 8, dtype=f64, device=dev)
    wide[:, ::2] = base
    pv = wide[:, ::2].detach().requires_grad_(True)
    (mops.circuit_columns(circ, pv) * w).sum().backward()
    np.testing.assert_allclose(pv.grad.cpu().numpy(), p64.grad.cpu().numpy(), rtol=1e-12, atol=1e-13)


@pytest.mark.gpu
def test_ops_trace_under_fake_tensor_mode():
    """make_fx in fake mode records the mcb200 operators without launching them (what torch.compile relies on)."""
    from torch.fx.experimental.proxy_tensor import make_fx

    dev = "cuda"
    circ = mops.CompiledCircuit(4, [mops.GateSpec(mops.GATE_RXRX, 0, 1, 0), mops.GateSpec(mops.GATE_FSWAP, 1, 2)], 2)
    h = circ.handle

    def f(p):
        x = torch.ops.mcb200.circuit_columns(p, h, None)
        cov = torch.ops.mcb200.cov_basis(x, 4, None)
        return torch.ops.mcb200.pfaffian(cov, False, 0.0625, 0.0)

    p = torch.rand(6, 2, dtype=f64, device=dev)
    gm = make_fx(f, tracing_mode="fake")(p)
    targets = [str(n.target) for n in gm.graph.nodes if n.op == "call_function"]
    assert targets == ["mcb200.circuit_columns.default", "mcb200.cov_basis.default", "mcb200.pfaffian.default"]
    torch.testing.assert_close(gm(p), f(p), rtol=0, atol=0)
