"""Generates tests/golden/*.json|npz from the REFERENCE ITSELF: the reference's own solver templates and CPU handlers
compiled in place from /root/reference (oracle/_ref/libitsolv_ref.so, built by oracle/Makefile). Run in the authoring
container (the GPU box has no /root/reference):   python tests/golden/make_golden.py
The fixtures pin (a) the C restatement oracle/itsolv_oracle.c and (b) the CUDA path, on boxes where the reference
build is not available."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

import itsolv_oracle_lib as ol  # noqa: E402
from iterative_solver_b200 import _native as N  # noqa: E402
from iterative_solver_b200 import harness as H  # noqa: E402

SOLVE_CASES = {
    # BASELINE.json configs[0]: examples/LinearEigensystemExample.cpp, ExampleProblem, 1 root
    "example_davidson_n20_r1": dict(n=20, kind=N.KIND_DAVIDSON, problem=N.PROBLEM_EXAMPLE, nroots=1, hermitian=0),
    "example_davidson_n20_r2": dict(n=20, kind=N.KIND_DAVIDSON, problem=N.PROBLEM_EXAMPLE, nroots=2, hermitian=0),
    "example_davidson_n20_r1_herm": dict(n=20, kind=N.KIND_DAVIDSON, problem=N.PROBLEM_EXAMPLE, nroots=1, hermitian=1),
    "example_davidson_n200_r4_herm": dict(n=200, kind=N.KIND_DAVIDSON, problem=N.PROBLEM_EXAMPLE, nroots=4, hermitian=1),
    "example_lineq_n20_r1": dict(n=20, kind=N.KIND_LINEQ, problem=N.PROBLEM_EXAMPLE, nroots=1, hermitian=1),
    "example_lineq_n200_r8": dict(n=200, kind=N.KIND_LINEQ, problem=N.PROBLEM_EXAMPLE, nroots=8, hermitian=1),
    "example_diis_n20": dict(n=20, kind=N.KIND_DIIS, problem=N.PROBLEM_EXAMPLE, max_size_qspace=6),
    "example_diis_n200": dict(n=200, kind=N.KIND_DIIS, problem=N.PROBLEM_EXAMPLE, max_size_qspace=6),
    # the synthetic banded operator of the benchmark at sizes the CPU finishes in well under a second
    "banded_davidson_n100000_r4": dict(n=100000, kind=N.KIND_DAVIDSON, nroots=4, hermitian=1),
    "banded_davidson_n20011_r4_nonherm": dict(n=20011, kind=N.KIND_DAVIDSON, nroots=4, hermitian=0),
    "banded_davidson_n30000_r16": dict(n=30000, kind=N.KIND_DAVIDSON, nroots=16, hermitian=1),
    "banded_davidson_n30000_r6_qcap8": dict(n=30000, kind=N.KIND_DAVIDSON, nroots=6, hermitian=1, max_size_qspace=8),
    "banded_davidson_n30000_r6_buf2": dict(n=30000, kind=N.KIND_DAVIDSON, nroots=6, nbuffers=2, hermitian=1),
    "banded_davidson_n30000_r4_p20": dict(n=30000, kind=N.KIND_DAVIDSON, nroots=4, hermitian=1, max_p=20),
    # P spaces too small to hold the solutions: P, Q (and D) together through several iterations
    "banded_davidson_n30000_r4_p4": dict(n=30000, kind=N.KIND_DAVIDSON, nroots=4, hermitian=1, max_p=4),
    "banded_davidson_n30000_r4_p6": dict(n=30000, kind=N.KIND_DAVIDSON, nroots=4, hermitian=1, max_p=6),
    "banded_davidson_n30000_r6_qcap8_buf3_p8": dict(n=30000, kind=N.KIND_DAVIDSON, nroots=6, nbuffers=3, hermitian=1,
                                                    max_size_qspace=8, max_p=8),
    "banded_davidson_n30000_r4_wide": dict(n=30000, kind=N.KIND_DAVIDSON, nroots=4, hermitian=1, half_bandwidth=16, eps=1e-2),
    # LinearEquations: right-hand sides b_k = A x_k with the scaled known solutions (ITSOLV_RHS_SCALED, the default)
    "banded_lineq_n50000_r1": dict(n=50000, kind=N.KIND_LINEQ, nroots=1, hermitian=1),
    "banded_lineq_n1000_r3": dict(n=1000, kind=N.KIND_LINEQ, nroots=3, hermitian=1),
    "banded_lineq_n100000_r3": dict(n=100000, kind=N.KIND_LINEQ, nroots=3, hermitian=1),
    "banded_lineq_n20000_r8": dict(n=20000, kind=N.KIND_LINEQ, nroots=8, hermitian=1),
    "banded_lineq_n20000_r8_qcap12": dict(n=20000, kind=N.KIND_LINEQ, nroots=8, hermitian=1, max_size_qspace=12),
    "banded_lineq_n30000_r8_buf3": dict(n=30000, kind=N.KIND_LINEQ, nroots=8, nbuffers=3, hermitian=1),
    "banded_lineq_n30000_r4_p20": dict(n=30000, kind=N.KIND_LINEQ, nroots=4, hermitian=1, max_p=20),
    "banded_lineq_n1000_r3_legacy_rhs": dict(n=1000, kind=N.KIND_LINEQ, nroots=3, hermitian=1, rhs_kind=N.RHS_LEGACY),
    "banded_lineq_n20000_r8_legacy_rhs": dict(n=20000, kind=N.KIND_LINEQ, nroots=8, hermitian=1, rhs_kind=N.RHS_LEGACY),
    # DIIS on r(v) = A (v - t): t(i) = 1/(i+1) by default, t = 1 for the legacy inputs
    "banded_diis_n50000": dict(n=50000, kind=N.KIND_DIIS, max_size_qspace=6),
    "banded_diis_n200000_qcap3": dict(n=200000, kind=N.KIND_DIIS, max_size_qspace=3),
    "banded_diis_n50000_legacy_target": dict(n=50000, kind=N.KIND_DIIS, max_size_qspace=6, rhs_kind=N.RHS_LEGACY),
}


def solve_record(ref, kw):
    spec = H.make_spec(trace=1, **kw)
    res, sol = ref.solve(spec, want_solutions=True)
    nroots = res.nroots
    tr = ref.read_trace()
    return {
        "spec": kw,
        "converged": int(res.converged),
        "iterations": int(res.iterations),
        "nwork_final": int(res.nwork_final),
        "eigenvalues": [float(res.eigenvalues[i]) for i in range(nroots)] if kw["kind"] == N.KIND_DAVIDSON else [],
        "errors": [float(res.errors[i]) for i in range(nroots)],
        "creations": [int(res.r_creations), int(res.q_creations), int(res.p_creations), int(res.d_creations)],
        "trace_shapes": [[op, r, c] for op, r, c, _ in tr],
        "trace_head": [v.tolist() for _, _, _, v in tr[:40]],
        "solution_checksums": [float(np.sum(s)) for s in sol],
        "solution_head": [s[:8].tolist() for s in sol],
    }


def handler_records(ref):
    """inputs are regenerated from the seed by the tests; outputs are stored"""
    out = {}
    rng = np.random.default_rng(2024)
    n = 4099
    X, Y = rng.standard_normal((4, n)), rng.standard_normal((6, n))
    alpha = rng.standard_normal((4, 6))
    shift = np.array([0.9, 1.9, 2.9, 3.9])
    diag = np.arange(1, n + 1, dtype=np.float64)
    out["seed"], out["n"] = 2024, n
    out["gemm_inner"] = ref.gemm_inner(X, Y)
    out["gemm_outer"] = ref.gemm_outer(alpha, X, Y)
    out["axpy"] = ref.axpy(0.37, X[0].copy(), Y[0].copy())
    out["scal"] = ref.scal(-1.7, X[1].copy())
    out["dot"] = np.array([ref.dot(X[0].copy(), Y[0].copy())])
    out["precondition"] = ref.precondition(X, shift, diag)
    xr = np.round(X[2], 1)
    for name, kw in (("select_min", {}), ("select_max", {"max": True}), ("select_absmax", {"max": True, "ignore_sign": True})):
        i, v = ref.select(xr.copy(), 25, **kw)
        out[name + "_idx"], out[name + "_val"] = i, v
    i, v = ref.select(xr.copy(), 25, y=np.round(Y[2], 1).copy())
    out["select_maxdot_idx"], out["select_maxdot_val"] = i, v
    V, nulls = ref.modified_gram_schmidt(np.vstack([X, X[0] + X[1]]), 1e-10)
    out["mgs"], out["mgs_nulls"] = V, np.array(nulls)
    maps = [{int(i): float(w) for i, w in zip(rng.choice(n, 3, replace=False), rng.standard_normal(3))} for _ in range(5)]
    out["sparse_gemm_inner"] = ref.sparse_gemm_inner(X, maps)
    out["sparse_gemm_outer"] = ref.sparse_gemm_outer(rng.standard_normal((5, 4)), maps, X)
    out["banded_apply"] = ref.banded_apply(X[3].copy(), 4, 1e-3)
    out["distribution_10_3"] = ref.distribution(10, 3)
    out["distribution_2e9_8"] = ref.distribution(2_000_000_001, 8)
    return out


def main():
    o = ol.load()
    if o.ref is None:
        raise SystemExit("oracle/_ref/libitsolv_ref.so is not built (needs /root/reference)")
    solves = {name: solve_record(o.ref, kw) for name, kw in SOLVE_CASES.items()}
    with open(os.path.join(HERE, "solve_golden.json"), "w") as f:
        json.dump(solves, f, indent=1)
    np.savez_compressed(os.path.join(HERE, "handler_golden.npz"), **handler_records(o.ref))
    for k, v in solves.items():
        print(k, "it", v["iterations"], "conv", v["converged"], v["eigenvalues"][:4], "%.2e" % max(v["errors"]))


if __name__ == "__main__":
    main()
