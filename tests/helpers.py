"""Test helpers: translate oracle Op lists into the device circuit descriptor."""
import numpy as np

from oracle.gates import Op


def ops_to_circuit(ops, n_qubits, batch):
    """oracle Op list -> (list[GateSpec], params (batch, P) float64, consts).  SPTM-flavoured and matchgate-flavoured
    rotations map to the same closed-form device gates; dense 4x4 ``Sptm`` ops become BLOCK4 gates."""
    from matchcake_b200 import ops as mops

    kind_map = {"CompRxRx": mops.GATE_RXRX, "SptmCompRxRx": mops.GATE_RXRX, "CompRyRy": mops.GATE_RYRY,
                "SptmCompRyRy": mops.GATE_RYRY, "CompRzRz": mops.GATE_RZRZ, "SptmCompRzRz": mops.GATE_RZRZ}
    gates, cols, consts = [], [], []
    off = 0
    for op in ops:
        w = op.wires
        if op.kind in kind_map:
            p = np.broadcast_to(np.asarray(op.params, dtype=np.float64), (batch, 2))
            gates.append(mops.GateSpec(kind_map[op.kind], w[0], w[1], off))
            cols.append(p)
            off += 2
        elif op.kind in ("fSWAP", "SptmFSwap"):
            lo, hi = min(w), max(w)
            gates.append(mops.GateSpec(mops.GATE_FSWAP, lo, hi, 0, mops.FLAG_TRANSPOSE if w[0] != lo else 0))
        elif op.kind == "Sptm":
            m = np.asarray(op.matrix, dtype=np.float64)
            assert m.shape[-1] == 4
            if m.ndim == 2:
                gates.append(mops.GateSpec(mops.GATE_BLOCK4_CONST, w[0], w[0] + 1, len(consts)))
                consts.extend(m.reshape(-1).tolist())
            else:
                gates.append(mops.GateSpec(mops.GATE_BLOCK4_PARAM, w[0], w[0] + 1, off))
                cols.append(m.reshape(batch, 16))
                off += 16
        else:
            raise ValueError(op.kind)
    params = np.concatenate(cols, axis=1) if cols else np.zeros((batch, 0))
    return gates, np.ascontiguousarray(params), (np.asarray(consts) if consts else None)


def readme_ops(n, params, matchgate=True):
    names = ["CompRxRx", "CompRyRy", "CompRzRz"] if matchgate else ["SptmCompRxRx", "SptmCompRyRy", "SptmCompRzRz"]
    ops = []
    for w in range(0, n - 1, 2):
        for k in names:
            ops.append(Op(k, (w, w + 1), params=np.asarray(params)))
    for w in range(1, n - 1, 2):
        ops.append(Op("fSWAP" if matchgate else "SptmFSwap", (w, w + 1)))
    return ops


def random_skew(rng, shape):
    a = rng.standard_normal(shape)
    return a - np.swapaxes(a, -1, -2)


# ------------------------------------------------------------------ dense state-vector simulation (2^n amplitudes)
# The reference's own tests compare the device with qml.device("default.qubit") on small circuits
# (tests/test_devices/test_nif_device/test_multiples_matchgate_circuit.py:66-102).  PennyLane is not
# installed here, so this is the same check by hand: apply the complex 4 x 4 matchgates (oracle.gates.Op.matchgate_matrix:
# comp_rotations.py:47-80, fermionic_swap.py:15-52) to the 2^n state vector, wire 0 = most significant bit
# (PennyLane's convention, the one states_to_binary follows, nif_device.py:151-165).
_SPTM_TO_MATCHGATE = {"SptmCompRxRx": "CompRxRx", "SptmCompRyRy": "CompRyRy", "SptmCompRzRz": "CompRzRz",
                      "SptmFSwap": "fSWAP"}


def statevector(ops, n_wires, init_bits=None):
    """U |init_bits> as a (2,) * n_wires array for a list of unbatched Ops on ADJACENT wire pairs.  SPTM-flavoured
    rotations / fSWAPs are simulated as the matchgates they are the single-particle picture of."""
    psi = np.zeros((2,) * n_wires, dtype=complex)
    psi[tuple(int(b) for b in (init_bits if init_bits is not None else [0] * n_wires))] = 1.0
    for op in ops:
        if op.kind in _SPTM_TO_MATCHGATE:
            op = Op(_SPTM_TO_MATCHGATE[op.kind], op.wires, params=op.params)
        w0, w1 = op.wires
        assert w1 == w0 + 1, "adjacent wires only"
        u = np.asarray(op.matchgate_matrix(), dtype=complex).reshape(2, 2, 2, 2)   # (out0, out1, in0, in1)
        psi = np.moveaxis(np.tensordot(u, psi, axes=([2, 3], [w0, w1])), (0, 1), (w0, w1))
    return psi


def statevector_probs(ops, n_wires, init_bits=None, wires=None):
    """Outcome probabilities of ``wires`` (default: all), shape (2^k,), MSB first in the order the wires are given."""
    p = np.abs(statevector(ops, n_wires, init_bits)) ** 2
    if wires is not None:
        keep = list(wires)
        p = p.sum(axis=tuple(i for i in range(n_wires) if i not in keep))
        p = np.transpose(p, np.argsort(np.argsort(keep)))  # axes in the order the wires were given
    return p.reshape(-1)
