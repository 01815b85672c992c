"""``FermionicPQCKernel`` (mirrors /root/reference/src/matchcake/ml/kernels/fermionic_pqc_kernel.py:28-199)."""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch

from ... import ops as _ops
from ...operations import SptmAngleEmbedding, SptmCompZX, SingleParticleTransitionMatrixOperation, matchgate_to_sptm_block
from .nif_kernel import NIFKernel

_ROT_KIND = {"X": _ops.GATE_RXRX, "Y": _ops.GATE_RYRY, "Z": _ops.GATE_RZRZ}


def pattern_double(n: int):
    return [(i, i + 1) for i in range(0, n - 1, 2)]


def pattern_double_odd(n: int):
    return [(i, i + 1) for i in range(1, n - 1, 2)]


def _hh_block() -> np.ndarray:
    h = np.array([[1, 1], [1, -1]], dtype=complex) / np.sqrt(2)
    m = np.zeros((4, 4), dtype=complex)
    m[0, 0], m[0, 3], m[3, 0], m[3, 3] = h[0, 0], h[0, 1], h[1, 0], h[1, 1]
    m[1, 1], m[1, 2], m[2, 1], m[2, 2] = h[0, 0], h[0, 1], h[1, 0], h[1, 1]
    return matchgate_to_sptm_block(m)


class FermionicPQCKernel(NIFKernel):
    DEFAULT_ROTATIONS = "Y,Z"
    DEFAULT_ENTANGLING_MTH = "fswap"
    available_entangling_mth = {"fswap", "identity", "hadamard"}

    def __init__(self, *, gram_batch_size: int = NIFKernel.DEFAULT_GRAM_BATCH_SIZE,
                 random_state: int = NIFKernel.DEFAULT_RANDOM_STATE, alignment: bool = NIFKernel.DEFAULT_ALIGNMENT,
                 alignment_iterations: int = NIFKernel.DEFAULT_ALIGNMENT_ITERATIONS,
                 alignment_learning_rate: float = NIFKernel.DEFAULT_ALIGNMENT_LEARNING_RATE,
                 alignment_early_stopping_patience: int = NIFKernel.DEFAULT_ALIGNMENT_EARLY_STOPPING_PATIENCE,
                 alignment_early_stopping_threshold: float = NIFKernel.DEFAULT_ALIGNMENT_EARLY_STOPPING_THRESHOLD,
                 n_qubits: int = NIFKernel.DEFAULT_N_QUBITS, r_dtype: Optional[torch.dtype] = None,
                 c_dtype: Optional[torch.dtype] = None, rotations: str = DEFAULT_ROTATIONS,
                 entangling_mth: str = DEFAULT_ENTANGLING_MTH):
        super().__init__(gram_batch_size=gram_batch_size, random_state=random_state, alignment=alignment,
                         alignment_iterations=alignment_iterations, alignment_learning_rate=alignment_learning_rate,
                         alignment_early_stopping_patience=alignment_early_stopping_patience,
                         alignment_early_stopping_threshold=alignment_early_stopping_threshold,
                         n_qubits=n_qubits, r_dtype=r_dtype, c_dtype=c_dtype)
        self.rotations = rotations
        self.entangling_mth = entangling_mth
        if self.entangling_mth not in self.available_entangling_mth:
            raise ValueError(f"Unknown entangling method: {self.entangling_mth}.")
        self.depth_: Optional[int] = None
        self.register_parameter("bias_", None)                # trainable (kernel alignment), float64 on the GPU
        self.data_scaling_: Optional[torch.Tensor] = None
        self._circuit = None
        self._circuit_key = None
        self._plan = None

    def fit(self, x_train, y_train=None):
        x_arr = x_train.detach().cpu().numpy() if isinstance(x_train, torch.Tensor) else np.asarray(x_train)
        n_inputs = int(np.prod(x_arr.shape[1:]))
        # fermionic_pqc_kernel.py:155-160: bias_ is the kernel's only registered parameter
        self.bias_ = torch.nn.Parameter(torch.from_numpy(self.np_rn_gen.random(n_inputs)).to("cuda"))
        self.data_scaling_ = torch.pi * torch.ones(n_inputs, dtype=torch.float64, device="cuda")
        if self.depth_ is None:
            self.depth_ = int(max(1, np.ceil(x_arr.shape[-1] / self.n_qubits)))
        self._circuit = None
        self._plan = None
        self._covs_train = None
        self._blocks_train = None
        super().fit(x_train, y_train)
        return self

    # ------------------------------------------------------------------ op-object view (API compatibility)
    def ansatz(self, x):
        """Yields the same operation stream as the reference (fermionic_pqc_kernel.py:166-199)."""
        xt = torch.as_tensor(np.asarray(x) if not isinstance(x, torch.Tensor) else x, dtype=torch.float64)
        xt = xt.reshape(len(xt), -1).cpu()
        xt = self.bias_.detach().cpu() + self.data_scaling_.cpu() * xt
        patterns = [pattern_double(self.n_qubits), pattern_double_odd(self.n_qubits)]
        for layer in range(self.depth_):
            sub_x = xt[..., layer * self.n_qubits:(layer + 1) * self.n_qubits]
            yield from SptmAngleEmbedding(sub_x, wires=self.wires, rotations=self.rotations).decomposition()
            for wires in patterns[layer % 2]:
                if self.entangling_mth == "fswap":
                    yield SptmCompZX(wires=wires)
                elif self.entangling_mth == "hadamard":
                    yield SingleParticleTransitionMatrixOperation(_hh_block(), wires=wires)
                elif self.entangling_mth != "identity":
                    raise ValueError(f"Unknown entangling method: {self.entangling_mth}")

    # ------------------------------------------------------------------ compiled view (what actually runs)
    def _compiled(self):
        """The ansatz as ONE device circuit; features are the parameter columns and ``bias + pi * x`` is applied
        inside the kernel, so no per-gate Python objects and no elementwise pre-pass exist on the hot path."""
        key = (self.bias_._version, self.bias_.data_ptr())  # an optimiser step on bias_ invalidates the baked affine
        if self._circuit is not None and self._circuit_key == key:
            return self._circuit
        self._circuit_key = key
        nq, f = self.n_qubits, len(self.bias_)
        gates: List[_ops.GateSpec] = []
        col_index: List[int] = []   # parameter column -> feature column (f = the constant-zero pad column)
        consts: List[float] = []
        patterns = [pattern_double(nq), pattern_double_odd(nq)]
        for layer in range(self.depth_):
            cols = list(range(layer * nq, min((layer + 1) * nq, f)))
            if len(cols) % 2:
                cols.append(f)  # SptmAngleEmbedding.pad_params: odd feature count -> zero angle
            for i in range(len(cols) // 2):
                off = len(col_index)
                col_index.extend([cols[2 * i], cols[2 * i + 1]])
                for r in self.rotations.split(","):
                    gates.append(_ops.GateSpec(_ROT_KIND[r], 2 * i, 2 * i + 1, off))
            for w0, w1 in patterns[layer % 2]:
                if self.entangling_mth == "fswap":
                    gates.append(_ops.GateSpec(_ops.GATE_FSWAP, w0, w1))
                elif self.entangling_mth == "hadamard":
                    gates.append(_ops.GateSpec(_ops.GATE_BLOCK4_CONST, w0, w1, len(consts)))
                    consts.extend(_hh_block().reshape(-1).tolist())
        idx = np.asarray(col_index, dtype=np.int64)
        bias = np.concatenate([self.bias_.detach().cpu().numpy(), [0.0]])[idx]
        scale = np.concatenate([self.data_scaling_.cpu().numpy(), [0.0]])[idx]
        identity = len(idx) == f and np.array_equal(idx, np.arange(f))
        circ = _ops.CompiledCircuit(nq, gates, len(idx), scale=scale, bias=bias,
                                    consts=np.asarray(consts) if consts else None)
        self._compiled_host = (gates, scale, bias, consts)
        self._circuit = (circ, None if identity else torch.as_tensor(idx, device="cuda"))
        self._circuit_plain = None
        self._gates_host = (gates, consts)
        return self._circuit

    def _compiled_plain(self) -> _ops.CompiledCircuit:
        """Same gate list WITHOUT the in-kernel affine map: under autograd ``bias_ + scale * x`` is a torch expression
        so that the chain rule reaches ``bias_`` (the adjoint kernel returns d/d angle)."""
        circ, _ = self._compiled()
        if self._circuit_plain is None:
            gates, consts = self._gates_host
            self._circuit_plain = _ops.CompiledCircuit(self.n_qubits, gates, circ.n_param_cols,
                                                       consts=np.asarray(consts) if consts else None)
        return self._circuit_plain

    def _features(self, x) -> torch.Tensor:
        xt = x if isinstance(x, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(x))
        xt = xt.to(device="cuda", dtype=torch.float64).reshape(len(xt), -1)
        circ, gather = self._compiled()
        if gather is not None:
            xt = torch.cat([xt, torch.zeros((len(xt), 1), dtype=torch.float64, device=xt.device)], dim=1)
            xt = xt.index_select(1, gather)
        return xt.contiguous()

    def _x_to_sptm(self, x) -> torch.Tensor:
        if self._differentiating():
            xt = x if isinstance(x, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(x))
            xt = xt.to(device="cuda", dtype=torch.float64).reshape(len(xt), -1)
            angles = self.bias_ + self.data_scaling_ * xt                  # fermionic_pqc_kernel.py:181
            _, gather = self._compiled()
            if gather is not None:
                angles = torch.cat([angles, torch.zeros((len(xt), 1), dtype=torch.float64, device="cuda")], dim=1)
                angles = angles.index_select(1, gather)
            return _ops.circuit_columns(self._compiled_plain(), angles.contiguous())
        return _ops.circuit_columns(self._compiled()[0], self._features(x))

    # ------------------------------------------------------------------ independent wire pairs (block-diagonal SPTM)
    def _block_plan(self):
        """If no gate couples two different wire pairs (depth_ = 1: rotations and entanglers act on the same pairs,
        fermionic_pqc_kernel.py:183-192) return one 2-qubit sub-circuit per pair; else None."""
        if self._plan is not None:
            return self._plan if self._plan is not False else None
        self._compiled()
        gates, scale, bias, consts = self._compiled_host
        parent = list(range(self.n_qubits))

        def find(a):
            while parent[a] != a:
                parent[a] = parent[parent[a]]
                a = parent[a]
            return a

        for g in gates:
            for w in range(g.wire0, g.wire1):
                parent[find(w)] = find(w + 1)
        comps = {}
        for g in gates:
            comps.setdefault(find(g.wire0), set()).update(range(g.wire0, g.wire1 + 1))
        comps = sorted(sorted(c) for c in comps.values())
        if len(comps) < 2 or any(len(c) != 2 or c[1] != c[0] + 1 for c in comps):
            self._plan = False
            return None
        self._plan = torch.as_tensor([c[0] for c in comps], dtype=torch.int32, device="cuda")
        return self._plan

    def _x_to_blocks(self, x) -> torch.Tensor:
        """(n_samples, n_pairs, 6): upper triangles of the 4x4 diagonal blocks of the per-sample covariances."""
        return _ops.circuit_pair_blocks(self._compiled()[0], self._features(x), self._block_plan())
