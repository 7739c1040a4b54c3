"""The drop-in boundary: both shared libraries load without a GPU and export every symbol the public headers declare;
compute calls fail loudly (no CPU fallback) when there is no device."""
import os
import re

import pytest

from iterative_solver_b200 import _native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(itsolv_[a-z0-9_]+)\s*\(", text)))


def test_kernel_library_exports_every_declared_symbol():
    lib = N.kernels()
    names = declared("itsolv_b200.h")
    assert len(names) >= 40
    for name in names:
        assert hasattr(lib, name), f"{name} is declared in include/itsolv_b200.h but not exported"
    assert sorted(N.KERNEL_API) == names, "the ctypes table must cover exactly the header"


def test_host_library_exports_every_declared_symbol():
    lib = N.host()
    names = declared("itsolv_b200_harness.h")
    for name in names:
        assert hasattr(lib, name), f"{name} is declared in include/itsolv_b200_harness.h but not exported"
    assert sorted(N.HARNESS_API) == names


def test_host_library_exports_the_flat_solver_interface():
    lib = N.host()
    text = open(os.path.join(ROOT, "include", "itsolv_b200_solver.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = sorted(set(re.findall(r"\b(ItsolvB200[A-Za-z]+)\s*\(", text)))
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} is declared in include/itsolv_b200_solver.h but not exported"
    assert sorted(N.SOLVER_API) == names
    # no instance is active: calls fail with a message instead of crashing (the reference throws, IterativeSolverCMPI.cpp:283)
    assert lib.ItsolvB200AddVector(1, None, None) == -1
    assert b"not initialised" in lib.ItsolvB200LastError()
    assert lib.ItsolvB200Finalize() != 0


@pytest.mark.parametrize("header", ["itsolv_b200.h", "itsolv_b200_harness.h", "itsolv_b200_solver.h"])
def test_public_headers_are_plain_c(header):
    """the boundary is a C ABI: every public header must compile as C99 (no C++ types, no torch types) and as C++"""
    import shutil
    import subprocess
    inc = os.path.join(ROOT, "include")
    for compiler, flags in (("gcc", ["-std=c99", "-x", "c"]), ("g++", ["-std=c++17", "-x", "c++"])):
        if shutil.which(compiler) is None:
            pytest.skip(f"{compiler} is not installed")
        r = subprocess.run([compiler, "-fsyntax-only", "-Wall", "-Wextra", "-pedantic", "-I" + inc] + flags +
                           [os.path.join(inc, header)], capture_output=True, text=True)
        assert r.returncode == 0 and not r.stderr.strip(), r.stderr


def test_struct_layouts_match_the_header():
    import ctypes as C
    assert C.sizeof(N.SolveSpec) == 80
    assert C.sizeof(N.SolveResult) == 16 + 2 * 64 * 8 + 3 * 8 + 4 * 8 + 7 * 8 + 2 * 8 + 8 + 7 * 8 + 2 * 8 + 4 * 8


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from iterative_solver_b200 import BackendError, Context
    with pytest.raises(BackendError):
        Context(0)
