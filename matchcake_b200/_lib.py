"""ctypes binding of libmcb200.so -- the C ABI declared in include/mcb200.h.

There is no CPU fallback: if the shared library is missing or was not built for this machine, importing the
compute layer raises.  ``load()`` is lazy so that pure host-side logic (circuit compilation, Jordan-Wigner
bookkeeping, Gram index rules) stays importable on a box without the library.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
# MCB200_LIB: another build of the same library (A/B measurements of compile-time variants)
LIB_PATH = os.environ.get("MCB200_LIB") or os.path.join(_HERE, "libmcb200.so")

_lib: Optional[C.CDLL] = None
_lock = threading.Lock()

c_i32, c_i64, c_f64, c_vp = C.c_int32, C.c_int64, C.c_double, C.c_void_p


class Mcb200Error(RuntimeError):
    """A libmcb200 entry point returned a non-zero status."""


class GateStruct(C.Structure):
    _fields_ = [("kind", c_i32), ("wire0", c_i32), ("wire1", c_i32), ("off", c_i32), ("flags", c_i32),
                ("reserved", c_i32)]


class CircuitStruct(C.Structure):
    _fields_ = [("gates", c_vp), ("layer_ptr", c_vp), ("n_gates", c_i32), ("n_layers", c_i32),
                ("n_qubits", c_i32), ("n_param_cols", c_i32), ("scale", c_vp), ("bias", c_vp), ("consts", c_vp),
                ("nearest_neighbour", c_i32), ("reserved", c_i32)]


# name -> (restype, argtypes); every symbol include/mcb200.h declares
SIGNATURES = {
    "mcb200_version": (c_i32, []),
    "mcb200_last_error": (C.c_char_p, []),
    "mcb200_device_count": (c_i32, []),
    "mcb200_set_abs_floor": (c_i32, [c_f64]),
    "mcb200_get_abs_floor": (c_f64, []),
    "mcb200_set_gram_pivot_growth": (c_i32, [c_f64]),
    "mcb200_get_gram_pivot_growth": (c_f64, []),
    "mcb200_set_expval_route": (c_i32, [c_i32]),
    "mcb200_circuit_validate": (c_i32, [C.POINTER(CircuitStruct)]),
    "mcb200_circuit_columns": (c_i32, [C.POINTER(CircuitStruct), c_vp, c_i64, c_vp, c_i32, c_vp, c_vp]),
    "mcb200_sptm_matmul": (c_i32, [c_vp, c_i64, c_i32, c_vp, c_i64, c_i32, c_vp, c_i64, c_i32, c_i64, c_vp]),
    "mcb200_matchgate_sptm": (c_i32, [c_vp, c_i64, c_vp, c_vp, c_vp]),
    "mcb200_product_state_cov": (c_i32, [c_vp, c_i64, c_i32, c_i32, c_vp, c_vp]),
    "mcb200_cov_basis": (c_i32, [c_vp, c_vp, c_i64, c_i32, c_i32, c_vp, c_vp]),
    "mcb200_circuit_cov_basis_workspace_bytes": (c_i64, [C.POINTER(CircuitStruct), c_i64, c_i32]),
    "mcb200_circuit_cov_basis": (c_i32, [C.POINTER(CircuitStruct), c_vp, c_i64, c_vp, c_i32, c_vp, c_vp, c_vp, c_vp]),
    "mcb200_cov_dense_workspace_bytes": (c_i64, [c_i64, c_i32, c_i32]),
    "mcb200_cov_dense": (c_i32, [c_vp, c_i64, c_vp, c_i64, c_i64, c_i32, c_i32, c_vp, c_i32, c_vp, c_vp, c_vp]),
    "mcb200_pfaffian_workspace_bytes": (c_i64, [c_i64, c_i32]),
    "mcb200_pfaffian": (c_i32, [c_vp, c_i64, c_i32, c_i32, c_f64, c_vp, c_vp, c_vp]),
    "mcb200_probs": (c_i32, [c_vp, c_vp, c_i64, c_i32, c_i32, c_i32, c_vp, c_vp]),
    "mcb200_probs_all_workspace_bytes": (c_i64, [c_i32]),
    "mcb200_probs_all": (c_i32, [c_vp, c_i64, c_i32, c_i32, c_vp, c_vp, c_vp]),
    "mcb200_sector_features": (c_i32, [c_vp, c_i64, c_i32, c_vp, c_i32, c_i32, c_vp, c_vp]),
    "mcb200_rowdot": (c_i32, [c_vp, c_vp, c_i64, c_i32, c_i32, c_vp, c_vp]),
    "mcb200_pauli_sum_small": (c_i32, [c_vp, c_i64, c_i32, c_vp, c_vp, c_vp, c_i32, c_i32, c_vp, c_vp]),
    "mcb200_circuit_expval_projector": (c_i32, [C.POINTER(CircuitStruct), c_vp, c_i64, c_vp, c_vp, c_vp, c_vp]),
    "mcb200_gram_workspace_bytes": (c_i64, [c_i64, c_i64, c_i32]),
    "mcb200_gram": (c_i32, [c_vp, c_vp, c_i64, c_i64, c_i32, c_i64, c_i64, c_i32, c_f64, c_vp, c_i64, c_vp, c_vp]),
    "mcb200_circuit_pair_blocks": (c_i32, [C.POINTER(CircuitStruct), c_vp, c_i64, c_vp, c_i32, c_vp, c_vp, c_vp]),
    "mcb200_gram_blocks4": (c_i32, [c_vp, c_vp, c_i64, c_i64, c_i32, c_i64, c_i64, c_i32, c_f64, c_vp, c_i64, c_vp]),
    "mcb200_gram_symmetrize": (c_i32, [c_vp, c_i64, c_i64, c_i64, c_vp]),
    "mcb200_sample": (c_i32, [c_vp, c_i64, c_i32, c_i64, C.c_uint64, c_vp, c_vp, c_vp, c_vp]),
    "mcb200_matchgate_sptm_backward": (c_i32, [c_vp, c_i64, c_vp, c_vp, c_vp]),
    "mcb200_pfaffian_backward": (c_i32, [c_vp, c_i64, c_i32, c_vp, c_vp, c_vp, c_vp]),
    "mcb200_probs_backward": (c_i32, [c_vp, c_vp, c_i64, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp]),
    "mcb200_sector_features_backward": (c_i32, [c_vp, c_i64, c_i32, c_vp, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp]),
    "mcb200_cov_basis_backward": (c_i32, [c_vp, c_vp, c_i64, c_i32, c_i32, c_vp, c_vp, c_vp]),
    "mcb200_cov_dense_backward_workspace_bytes": (c_i64, [c_i64, c_i32, c_i32]),
    "mcb200_cov_dense_backward": (c_i32, [c_vp, c_i64, c_vp, c_i64, c_i64, c_i32, c_i32, c_vp, c_i32, c_vp, c_vp, c_vp, c_vp]),
    "mcb200_circuit_columns_backward": (c_i32, [C.POINTER(CircuitStruct), c_vp, c_i64, c_i32, c_vp, c_vp, c_vp, c_vp]),
    "mcb200_launch_count": (c_i64, []),
    "mcb200_reset_launch_count": (None, []),
    "mcb200_probe_fp64_tflops": (c_f64, [c_i32, c_i32, c_vp]),
}


def load() -> C.CDLL:
    """Load libmcb200.so (once).  Raises if it is missing -- the product path has no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise Mcb200Error(
                f"{LIB_PATH} not found. Build it with `python -m matchcake_b200.build` (needs nvcc); "
                "matchcake_b200 has no CPU fallback."
            )
        lib = C.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().mcb200_last_error()
        raise Mcb200Error(msg.decode("utf-8", "replace") if msg else f"mcb200 error {rc}")


def launch_count() -> int:
    return int(load().mcb200_launch_count())


def set_expval_route(route: int) -> None:
    """0 = automatic, 1 = column route, 2 = covariance route of the fused projector kernel (sticky per thread)."""
    check(load().mcb200_set_expval_route(int(route)))


def reset_launch_count() -> None:
    load().mcb200_reset_launch_count()
