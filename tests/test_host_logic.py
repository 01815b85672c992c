"""CPU tests of the host-side logic of the product package (no CUDA compute)."""
import itertools
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from matchcake_b200 import distributed as dist_mod
from matchcake_b200 import observables as obs
from matchcake_b200 import ops as mops
from matchcake_b200.device import NonInteractingFermionicDevice, _PauliStructureCache
from matchcake_b200.ml.kernels import FermionicPQCKernel, GramMatrix
from matchcake_b200.operations import (CompRxRx, CompZX, MatchgateOperation, ProductState,
                                       SingleParticleTransitionMatrixOperation, SptmCompZX, matchgate_to_sptm_block)
from matchcake_b200.utils.jordan_wigner import JordanWigner, extend_majorana_indices
from oracle import covariance as ocv
from oracle import gates as og
from oracle import kernels as ok
from oracle import readout as oro

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_golden.json")))


def test_jordan_wigner_reference_table_bit_exact():
    jw = JordanWigner(3)
    for pauli, wires, idx, (re, im) in G["jordan_wigner_table"]["cases"]:
        mu, ph = jw.pauli_to_majorana(pauli, wires)
        assert mu.tolist() == idx and mu.dtype.kind == "i"
        assert ph == complex(re, im)


@pytest.mark.parametrize("n", [1, 2, 3, 4])
def test_jordan_wigner_all_words_match_oracle(n):
    jw = JordanWigner(n)
    for chars in itertools.product("IXYZ", repeat=n):
        w = "".join(chars)
        a, pa = jw.pauli_to_majorana(w, list(range(n)))
        b, pb = oro.pauli_to_majorana(w, list(range(n)))
        assert a.tolist() == b.tolist() and pa == pb
        ea, epa = extend_majorana_indices(a, pa, 2 * n)
        eb, epb = oro.extend_majorana_indices(b, pb, 2 * n)
        assert ea.tolist() == eb.tolist() and abs(epa - epb) < 1e-15


def test_jordan_wigner_long_random_words_and_wire_labels():
    rng = np.random.default_rng(0)
    n = 40
    jw = JordanWigner(n)
    for _ in range(200):
        w = "".join(rng.choice(list("IXYZ"), n, p=[0.7, 0.1, 0.1, 0.1]))
        a, pa = jw.pauli_to_majorana(w, list(range(n)))
        b, pb = oro.pauli_to_majorana(w, list(range(n)))
        assert a.tolist() == b.tolist() and pa == pb
    mu, ph = JordanWigner(2).pauli_to_majorana({"b": "X", "a": "Y"}, ["a", "b"])
    assert mu.tolist() == [0, 2] and ph == 1j  # == "YX"


def test_pauli_structure_cache_matches_oracle_payload():
    _PauliStructureCache.clear()
    words = ("XXI", "IYY", "ZZI", "IYX", "XYI", "IZI", "IIZ", "III", "XII")
    coeffs = np.arange(1, 10) * 0.1
    st = _PauliStructureCache.get(words, 3)
    ref = oro.pauli_sum_structure(words, coeffs, 3)
    assert set(st) == set(ref)
    for sector, (idx, phases, terms, wick) in st.items():
        np.testing.assert_array_equal(idx.reshape(len(terms), sector), ref[sector][0].reshape(len(terms), sector))
        np.testing.assert_allclose(np.real(coeffs[terms] * phases * wick), ref[sector][1], atol=1e-15)
    assert _PauliStructureCache.get(words, 3) is st  # cached


def test_gram_matrix_rules():
    """tests/test_ml/test_kernels/test_gram_matrix.py:40-122 of the reference."""
    for shape in [(10, 10), (12, 10), (10, 12), (4, 5), (5, 4), (1, 1), (3, 1), (1, 3)]:
        gm = GramMatrix(shape)
        got = np.concatenate(list(gm.indices_batch_generator(7)) or [np.zeros((0, 2), int)])
        ref = np.concatenate(list(ok.indices_batch_generator(shape, 7)) or [np.zeros((0, 2), int)])
        np.testing.assert_array_equal(got, ref)  # same cells, same order
        gm.apply_(lambda ids: -1.0)
        for i, j in np.ndindex(shape):
            assert gm[i, j].item() == (1.0 if i == j else -1.0)
        assert gm.to_tensor().dtype == torch.float32


def test_gram_matrix_memmap_store():
    """gram_matrix.py:25-47 of the reference: float32 entries in a numpy.memmap on a temporary file; same rules, the
    file holds what the container shows, and it is removed with the container."""
    import os

    shape = (9, 7)
    gm = GramMatrix(shape, memmap=True)
    path = gm._filepath
    assert os.path.exists(path) and os.path.getsize(path) == 9 * 7 * 4
    gm.apply_(lambda ids: (ids[:, 0] * 10 + ids[:, 1]).astype(np.float64), batch_size=5)
    ref = GramMatrix(shape).apply_(lambda ids: (ids[:, 0] * 10 + ids[:, 1]).astype(np.float64), batch_size=5)
    np.testing.assert_array_equal(gm.to_tensor().numpy(), ref.to_tensor().numpy())
    on_disk = np.fromfile(path, dtype=np.float32).reshape(shape)
    np.testing.assert_array_equal(on_disk, ref.to_tensor().numpy())
    # a finished Gram (here a host tensor standing in for the device one), streamed in row blocks
    k = torch.arange(63, dtype=torch.float64).reshape(shape) / 7
    g2 = GramMatrix.from_device(k, memmap=True, rows_per_copy=4)
    np.testing.assert_array_equal(np.fromfile(g2._filepath, dtype=np.float32).reshape(shape), k.to(torch.float32).numpy())
    p2 = g2._filepath
    g2.close()
    assert not os.path.exists(p2) and g2.to_tensor().shape == shape
    del gm
    assert not os.path.exists(path)


def test_states_to_binary_msb_first():
    got = NonInteractingFermionicDevice.states_to_binary(np.arange(8), 3)
    np.testing.assert_array_equal(got, oro.states_to_binary(np.arange(8), 3))
    assert got[3].tolist() == [0, 1, 1]


def test_layer_scheduler_preserves_order_and_disjointness():
    g = [mops.GateSpec(0, 0, 1), mops.GateSpec(1, 2, 3), mops.GateSpec(2, 0, 1), mops.GateSpec(3, 1, 4),
         mops.GateSpec(0, 5, 6), mops.GateSpec(0, 4, 5)]
    flat, ptr = mops.schedule_layers(g)
    assert flat == g and ptr == [0, 2, 3, 5, 6]
    for l in range(len(ptr) - 1):
        used = set()
        for gate in flat[ptr[l]:ptr[l + 1]]:
            w = set(range(gate.wire0, gate.wire1 + 1))
            assert not (used & w)
            used |= w


def test_matchgate_block_matches_trace_formula_and_validation():
    rng = np.random.default_rng(0)
    p = rng.uniform(0, 6, (5, 2))
    for cls, kind in [(CompRxRx, "CompRxRx")]:
        u = cls(p, wires=[0, 1]).matrix()
        np.testing.assert_allclose(u, og.Op(kind, (0, 1), params=p).matchgate_matrix(), atol=1e-15)
        np.testing.assert_allclose(matchgate_to_sptm_block(u), og.enforce_real(og.sptm_from_gate(u)), atol=1e-14)
    np.testing.assert_allclose(matchgate_to_sptm_block(CompZX(wires=[0, 1]).matrix()), og.sptm_fswap((0, 1)), atol=1e-15)
    with pytest.raises(ValueError):
        MatchgateOperation(np.ones((4, 4)), wires=[0, 1])
    with pytest.raises(AssertionError):
        MatchgateOperation(np.eye(4), wires=[0, 2])
    np.testing.assert_array_equal(SptmCompZX(wires=[3, 0]).matrix(), og.sptm_fswap((3, 0)))
    np.testing.assert_array_equal(SptmCompZX(wires=[0, 3]).matrix(), og.sptm_fswap((0, 3)))


def test_product_state_covariances_match_oracle():
    rng = np.random.default_rng(2)
    amp = rng.standard_normal((3, 5, 2)) + 1j * rng.standard_normal((3, 5, 2))
    amp /= np.linalg.norm(amp, axis=-1, keepdims=True)
    ps = ProductState(amp, wires=range(5))
    np.testing.assert_allclose(ps.covariance_matrix, ocv.covariance_matrix(amp), atol=1e-14)
    np.testing.assert_allclose(ps.displacement_vector, ocv.displacement_vector(amp), atol=1e-14)
    np.testing.assert_allclose(ps.extended_covariance_matrix,
                               ocv.extended_covariance_matrix(ocv.covariance_matrix(amp), ocv.displacement_vector(amp)), atol=1e-14)
    b = ProductState.from_basis_state([1, 0, 1], wires=[0, 1, 2])
    assert b.is_basis_state and b.basis_state_bits().tolist() == [1, 0, 1]
    with pytest.raises(ValueError):
        ProductState(np.ones((3, 2)), wires=range(3))  # not normalised
    with pytest.raises(ValueError):
        ProductState(np.ones(5), wires=range(3))


def test_sptm_op_validation():
    with pytest.raises(ValueError):
        SingleParticleTransitionMatrixOperation(np.eye(4) * (1 + 1j), wires=[0, 1])
    op = SingleParticleTransitionMatrixOperation(np.eye(4) + 1e-9j, wires=[0, 1])
    assert op.matrix().dtype == np.float64
    r = SingleParticleTransitionMatrixOperation.random(wires=[0, 1, 2], seed=0)
    np.testing.assert_allclose(r.matrix(), og.random_sptm(3, seed=0), atol=1e-14)


def test_device_construction_errors():
    with pytest.raises(AssertionError):
        NonInteractingFermionicDevice(wires=1)
    with pytest.raises(ValueError):
        NonInteractingFermionicDevice(wires=3, prob_strategy="nope")
    with pytest.raises(ValueError):
        NonInteractingFermionicDevice(wires=3, contraction_strategy="nope")
    assert NonInteractingFermionicDevice(wires=3, contraction_strategy=None).contraction_strategy == "none"
    with pytest.raises(ValueError):
        FermionicPQCKernel(entangling_mth="nope")


def test_observable_algebra():
    h = 0.1 * (obs.X(0) @ obs.X(1)) + 0.6 * (obs.Z(1) @ obs.I(2)) + obs.Y(2)
    coeffs, strings = obs.pauli_terms(h, (0, 1, 2))
    assert strings == ["XXI", "IZI", "IIY"] and coeffs == [0.1, 0.6, 1.0]
    with pytest.raises(ValueError):
        obs.pauli_terms(obs.X(7), (0, 1))


def test_balanced_row_ranges_cover_and_balance():
    for n0, n1, parts in [(8192, 8192, 8), (100, 100, 3), (262144, 4096, 8), (10, 12, 4), (5, 5, 8)]:
        rr = dist_mod.balanced_row_ranges(n0, n1, parts)
        assert rr[0][0] == 0 and rr[-1][1] == n0 and all(a[1] == b[0] for a, b in zip(rr, rr[1:]))
        cost = dist_mod.gram_row_cost(n0, n1)
        per = [cost[s:e].sum() for s, e in rr]
        if n0 >= 100:
            assert max(per) <= 1.05 * sum(per) / parts + n1
    assert dist_mod.even_ranges(10, 3) == [(0, 4), (4, 7), (7, 10)]
    # folded blocks: a partition of the rows, equal row counts, equal cell counts
    for n, parts in [(8192, 8), (8192, 4), (101, 3), (10, 4), (5, 8), (2, 2)]:
        blocks = dist_mod.folded_row_blocks(n, parts)
        rows = sorted(r for (t, b) in blocks for blk in (t, b) for r in range(*blk))
        assert rows == list(range(n))
        cnt = [(t[1] - t[0]) + (b[1] - b[0]) for t, b in blocks]
        assert max(cnt) == -(-n // parts)
        if n >= 100:
            cost = [sum(range(*t)) + sum(range(*b)) for t, b in blocks]
            assert max(cost) - min(c for c in cost if c) <= 2 * n


_GLOO_WORKER = r"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["REPO_ROOT"])
from matchcake_b200 import distributed as dm
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + os.environ["PORT"], rank=int(os.environ["RANK"]), world_size=2)
n0, n1 = 11, 7
full = torch.arange(n0 * n1, dtype=torch.float64).reshape(n0, n1)
rr = dm.balanced_row_ranges(n0, n1, 2)
out = dm.sharded_rows(lambda s, e: full[s:e].clone(), rr, n0)
assert torch.equal(out, full), out
ev = dm.even_ranges(9, 2)
vec = torch.arange(9, dtype=torch.float64) * 2
got = dm.sharded_rows(lambda s, e: vec[s:e].clone(), ev, 9)
assert torch.equal(got, vec)
for n in (2, 7, 10, 33):
    sq = torch.arange(n * n, dtype=torch.float64).reshape(n, n)
    got = dm.sharded_folded_rows(lambda s, e: sq[s:e].clone(), n)
    assert torch.equal(got, sq), (n, got)
# in-place folded gather: both halves land in their final rows (top over WORLD, bottom over the reversed group)
for n in (4, 8, 32):
    sq = torch.arange(n * n, dtype=torch.float64).reshape(n, n)
    def rows(s, e, out):
        out.copy_(sq[s:e])
    got = dm.folded_rows_in_place(rows, n, n, torch.float64, "cpu")
    assert torch.equal(got, sq), (n, got)
# uneven contiguous row blocks gathered into the final matrix
for n0, n1 in ((11, 7), (64, 5)):
    fullm = torch.arange(n0 * n1, dtype=torch.float64).reshape(n0, n1)
    rr = dm.balanced_row_ranges(n0, n1, 2)
    k = torch.full((n0, n1), -1.0, dtype=torch.float64)
    s, e = rr[dist.get_rank()]
    k[s:e] = fullm[s:e]
    assert torch.equal(dm.gather_row_ranges(k, rr), fullm)
dist.barrier(); dist.destroy_process_group()
print("ok", os.environ["RANK"])
"""


def test_all_gather_rows_world_size_2_gloo(tmp_path):
    """The one exchange step of the path (row-block all-gather), world_size 2 on the gloo backend."""
    script = tmp_path / "w.py"
    script.write_text(_GLOO_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                              env={**os.environ, "RANK": str(r), "PORT": port, "REPO_ROOT": ROOT})
             for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("ok" in o for o in outs)


def test_strategy_plugin_surface_contract():
    """The reference's strategy ABCs (SURVEY 8b iii): names, required kwargs and error behaviour, no GPU needed."""
    from matchcake_b200 import strategies as st
    from matchcake_b200.operations import BasisState, ProductState

    p = st.B200ProbabilityStrategy()
    assert p.NAME == "B200" and p.REQUIRES_KWARGS == ["global_sptm", "all_wires"]
    assert p.can_execute(BasisState(np.zeros(3, dtype=int), wires=range(3)))
    assert p.can_execute(ProductState(np.tile(np.array([1.0, 0.0], dtype=complex), (3, 1)), wires=range(3)))
    assert not p.can_execute(object())
    with pytest.raises(ValueError, match="Missing required keyword argument: global_sptm"):
        p(state_prep_op=None, target_binary_states=[0, 0, 0], wires=[0, 1, 2], all_wires=[0, 1, 2])
    with pytest.raises(ValueError, match="Missing required keyword argument"):
        st.B200PfaffianExpvalStrategy()(None, None, global_sptm=np.eye(4))
    assert isinstance(st.get_probability_strategy("LookupTable"), st.B200ProbabilityStrategy)
    assert isinstance(st.get_probability_strategy(p), st.B200ProbabilityStrategy)
    with pytest.raises(ValueError):
        st.get_probability_strategy("ExplicitSum")
    c = st.B200ContractionStrategy()
    assert c.get_next_operations("op-a") == [] and c.get_next_operations("op-b") == []
    assert c._ops == ["op-a", "op-b"]
    c.reset()
    assert c._ops == [] and c.get_reminding() is None
    for base in (st.ProbabilityStrategy, st.ExpvalStrategy):
        with pytest.raises(NotImplementedError):
            base()(state_prep_op=None, target_binary_states=None, wires=None) if base is st.ProbabilityStrategy else base()(None, None)


def test_gram_pivoting_option_is_parsed_on_the_host():
    """ops.gram(pivoting=...): the named bounds, an explicit number and the error for anything else are resolved without a
    launch; "auto" is strict for small problems and for d > 64 (whose kernels use guarded scalar pivots)."""
    import torch

    from matchcake_b200 import ops as mops

    assert mops._gram_growth_for(None, None, 0, "fast") == mops.GRAM_GROWTH_FAST == 1e5
    assert mops._gram_growth_for(None, None, 0, "strict") == mops.GRAM_GROWTH_STRICT == 1e3
    assert mops._gram_growth_for(None, None, 0, 250) == 250.0
    with pytest.raises(ValueError, match="pivoting must be"):
        mops._gram_growth_for(None, None, 0, "sometimes")
    small = torch.zeros((10, 64, 64), dtype=torch.float64)
    assert mops._gram_growth_for(small, small, 100, "auto") == mops.GRAM_GROWTH_STRICT       # 100 cells: no decision to make
    wide = torch.zeros((2, 128, 128), dtype=torch.float64)
    assert mops._gram_growth_for(wide, wide, 1 << 20, "auto") == mops.GRAM_GROWTH_STRICT     # d > 64
