"""CPU-side checks of the drop-in boundary: libmcb200.so loads and exports exactly what include/mcb200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "mcb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mcb200_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_loads_and_exports_every_declared_symbol():
    from matchcake_b200 import _lib, build

    build.build()
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 18
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/mcb200.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == names
    assert lib.mcb200_version() == 100
    assert lib.mcb200_launch_count() >= 0


def test_no_cpu_fallback_without_gpu():
    import torch

    from matchcake_b200 import ops as mops
    from matchcake_b200._lib import Mcb200Error

    with pytest.raises(Mcb200Error):
        mops.pfaffian(torch.zeros((1, 2, 2), dtype=torch.float64))


def test_product_never_imports_oracle():
    """The product package must not reference the oracle or the reference tree at run time (tier rule 3)."""
    pkg = os.path.join(ROOT, "matchcake_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert not re.search(r"(open|listdir|sys\.path[^\n]*)\([^\n]*/root/reference", txt), f


def test_no_batched_memcpy_calls_in_sources():
    """The profiling guide forbids the batched memcpy entry points on this driver; keep them out of the tree."""
    banned = ["cudaMemcpy" + "BatchAsync", "cudaMemcpy3D" + "BatchAsync", "cuMemcpy" + "BatchAsync", "cuMemcpy3D" + "BatchAsync"]
    for dirpath, dirs, files in os.walk(ROOT):
        dirs[:] = [d for d in dirs if d not in (".git", "gpurun_out", "build", "__pycache__")]
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".sh")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                for b in banned:
                    assert b not in txt, (f, b)


def _compile_c_client(out_path):
    """gcc (plain C99, no C++ and no Python) against include/mcb200.h, libmcb200.so and the CUDA runtime."""
    import shutil
    import subprocess

    from matchcake_b200 import build

    build.build()
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    if shutil.which("gcc") is None or not os.path.exists(os.path.join(cuda, "include", "cuda_runtime_api.h")):
        pytest.skip("gcc or the CUDA runtime headers are not available")
    libdir = os.path.join(ROOT, "matchcake_b200")
    cmd = ["gcc", "-std=c99", "-Wall", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(cuda, "include"),
           os.path.join(ROOT, "tests", "c", "abi_client.c"), "-o", out_path, "-L", libdir, "-lmcb200",
           "-L", os.path.join(cuda, "lib64"), "-lcudart", "-lm", f"-Wl,-rpath,{libdir}"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return out_path


def test_header_is_plain_c_and_a_c_client_links(tmp_path):
    """The boundary is a C ABI: the header must compile as C99 (no C++-isms) and a C program must link against it."""
    import subprocess

    src = tmp_path / "hdr.c"
    src.write_text('#include "mcb200.h"\nint main(void) { return mcb200_version() < 0; }\n')
    res = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only",
                          "-I", os.path.join(ROOT, "include"), str(src)], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    _compile_c_client(str(tmp_path / "abi_client"))


def test_integration_doc_lists_every_entry_point():
    """INTEGRATION.md is the maintainer-facing description of the boundary: it must name every exported symbol."""
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    missing = [n for n in declared_symbols() if n not in doc and not n.endswith("_workspace_bytes")]
    assert not missing, missing
