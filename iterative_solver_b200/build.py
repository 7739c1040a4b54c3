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


def build_host(force: bool = False) -> str:
    """The reference-facing plugin + harness. Needs the reference headers; otherwise the prebuilt library is kept."""
    lib = os.path.join(LIBDIR, "libitsolv_b200_host.so")
    if not os.path.isdir(os.path.join(REFERENCE, "src", "molpro")):
        if not os.path.exists(lib):
            raise RuntimeError("libitsolv_b200_host.so is not built and the reference headers are not available")
        return lib
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
    build_all(force="--force" in sys.argv)
    print("built", os.listdir(LIBDIR))
