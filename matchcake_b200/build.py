"""Builds libmcb200.so (hand-written sm_100a CUDA + the C ABI of include/mcb200.h) in-tree with nvcc.

Usage: ``python -m matchcake_b200.build [--force]``.  The shared library is git-ignored but travels to the GPU
box with the repository snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmcb200.so")
SOURCES = ["api.cu", "circuit.cu", "circuit_brick.cu", "circuit_panel.cu", "pfaffian.cu", "pfaffian_tiles.cu", "pfaffian_block_tiles.cu", "pfaffian_panel.cu", "expval_warp.cu", "expval_cov.cu", "gram_blocks.cu", "prep.cu", "gemm.cu", "backward.cu", "sampling.cu"]
HEADERS = ["common.cuh", "circuit_dev.cuh", "pfaffian_src.cuh", "pfaffian_dev.cuh", "pfaffian_tiles.cuh", "pfaffian_flat.cuh", "pfaffian_block_tiles.cuh", "tiles_api.cuh", os.path.join("..", "..", "include", "mcb200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v",
]


# extra nvcc flags for experiments, e.g. MCB200_NVCC_EXTRA="-DMCB200_PANEL_PROFILE" python -m matchcake_b200.build --force
NVCC_FLAGS += os.environ.get("MCB200_NVCC_EXTRA", "").split()


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: libmcb200.so cannot be built (there is no CPU fallback)")
    return nvcc


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(p) > t for p in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    nvcc = _nvcc()
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    for s in SOURCES:
        o = os.path.join(objdir, s.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, s), "-o", o]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(o)
    log = []
    for s, p in procs:
        out, _ = p.communicate()
        log.append(f"==== {s}\n{out}")
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {s}:\n{out}")
    with open(os.path.join(objdir, "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    cmd = [nvcc, "-shared", "-o", LIB, *objs, "-lcudart"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    if verbose:
        print("\n".join(log))
    return LIB


if __name__ == "__main__":
    try:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    except RuntimeError as e:  # make a failed build impossible to miss behind `| tail`
        print(str(e)[-3000:])
        print("BUILD FAILED")
        sys.exit(1)
