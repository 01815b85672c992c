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
