#!/usr/bin/env python
"""Runs one Davidson configuration through the three driver paths (reference's solver class on the CUDA handlers, batched
pieces under the reference's solve() loop, fused solve()) and prints what the parity tests compare."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import iterative_solver_b200 as pkg  # noqa: E402
from iterative_solver_b200 import _native as N  # noqa: E402
from iterative_solver_b200 import harness as H  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=30000)
ap.add_argument("--roots", type=int, default=16)
ap.add_argument("--qspace", type=int, default=8)
ap.add_argument("--buffers", type=int, default=8)
a = ap.parse_args()
ctx = pkg.Context(0)
ctx.init_comm(0, 1, b"\0" * 128)
for fused in (0, 2, 1):
    spec = H.make_spec(a.n, kind=N.KIND_DAVIDSON, nroots=a.roots, hermitian=1, max_size_qspace=a.qspace,
                       nbuffers=a.buffers, fused=fused)
    res, _ = H.solve(ctx, spec)
    print("fused", fused, "iterations", res.iterations, "converged", res.converged,
          "creations", [res.r_creations, res.q_creations, res.p_creations, res.d_creations],
          "max error %.3e" % max(res.errors[i] for i in range(a.roots)),
          "eig %.12f %.12f" % (res.eigenvalues[0], res.eigenvalues[a.roots - 1]), "launches", res.kernel_launches, flush=True)
ctx.close()
