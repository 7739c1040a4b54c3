"""Build the native libraries in-tree.

  libitsolv_b200.so       hand-written sm_100a kernels + the C ABI of include/itsolv_b200.h   (nvcc)
  libitsolv_b200_host.so  DistrArrayCUDA / ArrayHandlerCUDA plugged into the reference's solver templates, and the
                          solve harness of include/itsolv_b200_harness.h                      (g++, needs /root/reference)

The second library is compiled against the reference's own headers (it is a plugin for that API); on a box without
/root/reference the prebuilt file is used as is.
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "iterative_solver_b200")
CSRC = os.path.join(PKG, "csrc")
LIBDIR = os.path.join(PKG, "lib")
OBJDIR = os.path.join(PKG, "build")
REFERENCE = os.environ.get("ITSOLV_REFERENCE", "/root/reference")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
# The system compiler, whatever CXX says: a toolchain that links libstdc++ statically (this image exports
# CXX=/opt/gcc/bin/g++, which does) would put a second copy of the iostream/locale state into the library, and with
# -Bsymbolic (below) the library would use that copy while the process uses libstdc++.so's: a crash in the first
# operator<<(double). ITSOLV_CXX overrides.
CXX = os.environ.get("ITSOLV_CXX", "/usr/bin/g++")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-I" + os.path.join(ROOT, "include"), "--fmad=true",
]


def _run(cmd: list[str]) -> None:
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError("build step failed: " + cmd[0])


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build_kernels(force: bool = False, verbose_ptxas: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    sources = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    headers = glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(ROOT, "include", "*.h"))
    objs = []
    jobs = []
    for src in sources:
        obj = os.path.join(OBJDIR, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if force or _stale(obj, [src] + headers):
            cmd = [NVCC] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose_ptxas else []) + ["-c", src, "-o", obj]
            jobs.append(cmd)
    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        list(ex.map(_run, jobs))
    lib = os.path.join(LIBDIR, "libitsolv_b200.so")
    if force or jobs or _stale(lib, objs):
        _run([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + objs +
             ["-cudart", "static", "-ldl", "-lpthread", "-lrt"])
    return lib


def _openblas() -> str:
    import scipy
    cands = glob.glob(os.path.join(os.path.dirname(scipy.__file__), "..", "scipy.libs", "libscipy_openblas*.so"))
    if not cands:
        raise RuntimeError("scipy's bundled OpenBLAS (LAPACK for the host subspace algebra) was not found")
    return os.path.abspath(cands[0])


MANIFEST = os.path.join(PKG, "reference_manifest.json")


def reference_files() -> dict:
    """sha256 of every file of the reference checkout the host library is built from: the solver templates and array
    headers it instantiates and the seven translation units it compiles (paths relative to the checkout)."""
    import hashlib
    base = os.path.join(REFERENCE, "src", "molpro", "linalg")
    out = {}
    for sub, pats in (("", ("*.h", "options.cpp")), ("itsolv", ("**/*.h", "*.cpp")), ("array", ("**/*.h",))):
        for pat in pats:
            for path in sorted(glob.glob(os.path.join(base, sub, pat), recursive=True)):
                with open(path, "rb") as f:
                    out[os.path.relpath(path, REFERENCE)] = hashlib.sha256(f.read()).hexdigest()
    return out


def check_reference() -> None:
    """The solver layer is the reference's own (header templates, instantiated with the CUDA containers): the checkout is a
    build dependency, pinned by content in reference_manifest.json. A different revision is refused unless
    ITSOLV_REFERENCE_UNPINNED=1 (the template interfaces the plugin classes implement may have moved)."""
    import json
    if not os.path.exists(MANIFEST):
        return
    pinned = json.load(open(MANIFEST))["files"]
    found = reference_files()
    wrong = sorted(k for k in pinned if found.get(k) != pinned[k])
    if wrong and os.environ.get("ITSOLV_REFERENCE_UNPINNED", "0") != "1":
        raise RuntimeError(
            f"the reference checkout at {REFERENCE} is not the revision this package is pinned to: {len(wrong)} of "
            f"{len(pinned)} files differ or are missing (first: {wrong[0]}). Point ITSOLV_REFERENCE at a matching checkout of "
            "knowles-group/iterative-solver, re-pin with `python -m iterative_solver_b200.build --pin` after reviewing the "
            "plugin classes against the new headers, or set ITSOLV_REFERENCE_UNPINNED=1")


def pin_reference() -> None:
    import json
    with open(MANIFEST, "w") as f:
        json.dump({"what": "content pin of the knowles-group/iterative-solver files libitsolv_b200_host.so is built from "
                           "(the checkout carries no version tag)", "files": reference_files()}, f, indent=0, sort_keys=True)


def build_host(force: bool = False) -> str:
    """The reference-facing plugin + harness: the reference's solver templates instantiated with the CUDA containers.
    The reference checkout (ITSOLV_REFERENCE, default /root/reference) is a build dependency; a box without it (the GPU
    box) keeps the prebuilt library that travels with the tree."""
    lib = os.path.join(LIBDIR, "libitsolv_b200_host.so")
    if not os.path.isdir(os.path.join(REFERENCE, "src", "molpro")):
        if not os.path.exists(lib):
            raise RuntimeError(
                "libitsolv_b200_host.so is not built and cannot be: it instantiates the solver templates of "
                f"knowles-group/iterative-solver, and no checkout was found at {REFERENCE} (set ITSOLV_REFERENCE). The "
                "kernel library libitsolv_b200.so (include/itsolv_b200.h) does not need it.")
        return lib
    check_reference()
    os.makedirs(OBJDIR, exist_ok=True)
    refsrc = os.path.join(REFERENCE, "src", "molpro", "linalg")
    ref_cpp = [os.path.join(refsrc, p) for p in (
        "options.cpp", "itsolv/Logger.cpp", "itsolv/util.cpp", "itsolv/Options.cpp",
        "itsolv/LinearEigensystemDavidsonOptions.cpp", "itsolv/LinearEquationsDavidsonOptions.cpp",
        "itsolv/NonLinearEquationsDIISOptions.cpp")]
    own_cpp = [os.path.join(PKG, "host", "helper_lapack.cpp"), os.path.join(PKG, "harness", "solver_capi.cpp"),
               os.path.join(PKG, "harness", "solver_flat_capi.cpp")]
    headers = (glob.glob(os.path.join(PKG, "host", "*.h")) + glob.glob(os.path.join(PKG, "harness", "*.h")) +
               glob.glob(os.path.join(ROOT, "include", "*.h")))
    flags = ["-std=c++17", "-O2", "-DNDEBUG", "-fPIC", "-ffp-contract=off", "-w",
             "-I" + os.path.join(PKG, "host", "shim"), "-I" + os.path.join(PKG, "host"),
             "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(REFERENCE, "src")]
    objs, jobs = [], []
    for src in ref_cpp + own_cpp:
        obj = os.path.join(OBJDIR, "host_" + os.path.basename(src)[:-4] + ".o")
        objs.append(obj)
        if force or _stale(obj, [src] + headers):
            jobs.append([CXX] + flags + ["-c", src, "-o", obj])
    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        list(ex.map(_run, jobs))
    blas = _openblas()
    if force or jobs or _stale(lib, objs):
        # -Bsymbolic: the library's own references to the reference's template instantiations (eigenproblem, svd_system,
        # ...) bind to ITS definitions even when another library with the same symbols (the oracle build of the
        # reference, tests only) is loaded into the same process
        _run([CXX, "-shared", "-Wl,-Bsymbolic", "-o", lib] + objs +
             ["-L" + LIBDIR, "-litsolv_b200", "-Wl,-rpath,$ORIGIN", blas, "-Wl,-rpath," + os.path.dirname(blas)])
    return lib


def build_all(force: bool = False) -> None:
    build_kernels(force)
    build_host(force)


if __name__ == "__main__":
    if "--pin" in sys.argv:
        pin_reference()
        print("pinned", len(reference_files()), "reference files in", MANIFEST)
        sys.exit(0)
    build_all(force="--force" in sys.argv)
    print("built", os.listdir(LIBDIR))
