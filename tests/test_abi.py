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


def test_struct_layouts_match_the_header(tmp_path):
    """sizes and the offsets of the last members as the C compiler lays the structs of the header out"""
    import ctypes as C
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc is not installed")
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "itsolv_b200_harness.h"\n#include "itsolv_b200.h"\n'
                   'int main(void){printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(itsolv_solve_spec), '
                   'offsetof(itsolv_solve_spec, rhs_kind), sizeof(itsolv_solve_result), '
                   'offsetof(itsolv_solve_result, calls_residual), sizeof(itsolv_trace_entry), sizeof(itsolv_counters));return 0;}\n')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-I" + os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(t) for t in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert got == [C.sizeof(N.SolveSpec), N.SolveSpec.rhs_kind.offset, C.sizeof(N.SolveResult),
                   N.SolveResult.calls_residual.offset, C.sizeof(N.TraceEntry), C.sizeof(N.Counters)]


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from iterative_solver_b200 import BackendError, Context
    with pytest.raises(BackendError):
        Context(0)


def test_the_reference_checkout_is_pinned_by_content(tmp_path, monkeypatch):
    """The host library instantiates the reference's templates: build.py refuses a checkout that differs from
    reference_manifest.json and says what to do (an explicit, versioned build dependency)."""
    import json
    import shutil

    from iterative_solver_b200 import build as B
    pinned = json.load(open(B.MANIFEST))["files"]
    assert len(pinned) > 50 and all(len(v) == 64 for v in pinned.values())
    if not os.path.isdir(os.path.join(B.REFERENCE, "src", "molpro")):
        pytest.skip("no reference checkout on this box (the prebuilt library travels with the tree)")
    assert B.reference_files() == pinned
    B.check_reference()
    # a checkout with one header changed
    other = tmp_path / "reference"
    shutil.copytree(os.path.join(B.REFERENCE, "src", "molpro", "linalg"), other / "src" / "molpro" / "linalg")
    victim = other / "src" / "molpro" / "linalg" / "itsolv" / "IterativeSolver.h"
    victim.write_text(victim.read_text() + "\n// changed\n")
    monkeypatch.setattr(B, "REFERENCE", str(other))
    with pytest.raises(RuntimeError, match="not the revision this package is pinned to"):
        B.check_reference()
    monkeypatch.setenv("ITSOLV_REFERENCE_UNPINNED", "1")
    B.check_reference()
