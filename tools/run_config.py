#!/usr/bin/env python
"""Runs one of BASELINE.json's configurations on the CUDA backend and prints one JSON line (rank 0).
    python tools/run_config.py --config c3 [--scale 0.1]                       (1 GPU)
    torchrun --nproc-per-node 8 tools/run_config.py --config c4 [--scale 0.1]  (row-sharded)
configs: c2 Davidson n=1e7 4 roots | c3 LinearEquations n=2e8 8 RHS | c4 Davidson n=2e9 16 roots |
         c5a P-space (500) Davidson n=5e8 4 roots | c5b DIIS n=5e8"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import iterative_solver_b200 as pkg  # noqa: E402
from iterative_solver_b200 import _native as N  # noqa: E402
from iterative_solver_b200 import distributed as D  # noqa: E402
from iterative_solver_b200 import harness as H  # noqa: E402

CONFIGS = {
    "c2": dict(n=10_000_000, kind=N.KIND_DAVIDSON, nroots=4, hermitian=1),
    "c3": dict(n=200_000_000, kind=N.KIND_LINEQ, nroots=8, hermitian=1, max_size_qspace=24),
    "c4": dict(n=2_000_000_000, kind=N.KIND_DAVIDSON, nroots=16, hermitian=1, max_size_qspace=8),
    "c5a": dict(n=500_000_000, kind=N.KIND_DAVIDSON, nroots=4, hermitian=1, max_p=500),
    "c5b": dict(n=500_000_000, kind=N.KIND_DIIS, max_size_qspace=8),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", required=True, choices=sorted(CONFIGS))
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--qspace", type=int, default=-1)
    ap.add_argument("--buffers", type=int, default=0)
    ap.add_argument("--max-iter", type=int, default=0)
    ap.add_argument("--threshold", type=float, default=0.0)
    ap.add_argument("--unfused", action="store_true", help="the reference's solver class call for call (default: fused driver)")
    ap.add_argument("--fused-equations", action="store_true", help="LinearEquations / DIIS on the fused X space")
    ap.add_argument("--cold", action="store_true", help="report the first solve (includes first-touch allocation of the pool)")
    args = ap.parse_args()
    rank, world, local = D.env_rank_world()
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = pkg.Context(local)
    D.attach_communicator(ctx)
    kw = dict(CONFIGS[args.config])
    kw["n"] = max(1000, int(kw["n"] * args.scale))
    if args.qspace >= 0:
        kw["max_size_qspace"] = args.qspace
    if args.buffers:
        kw["nbuffers"] = args.buffers
    if args.max_iter:
        kw["max_iter"] = args.max_iter
    if args.threshold:
        kw["convergence_threshold"] = args.threshold
    # Davidson: the fused driver is the default. The equation solvers run the reference's classes unless --fused-equations
    # asks for their fused X space (FusedEquations.h).
    if (kw["kind"] == N.KIND_DAVIDSON and not args.unfused) or args.fused_equations:
        kw["fused"] = 1
    spec = H.make_spec(**kw)
    ctx.set_profiling(True)
    ctx.mem_usage(reset_peak=True)
    t0 = time.perf_counter()
    problem = H.Problem(ctx, spec)
    res = problem.solve(spec)
    if not args.cold:  # the stream-ordered pool is now populated: time the steady state
        t0 = time.perf_counter()
        res = problem.solve(spec)
    wall = time.perf_counter() - t0
    problem.close()
    live, peak = ctx.mem_usage()
    nloc = int(np.diff(pkg.distribution(spec.n, world))[rank])
    nroots = 1 if spec.kind == N.KIND_DIIS else spec.nroots
    line = {
        "config": args.config, "spec": kw, "n_gpus": world, "n_local": nloc, "converged": int(res.converged),
        "iterations": int(res.iterations), "errors": [res.errors[i] for i in range(nroots)],
        "eigenvalues": [res.eigenvalues[i] for i in range(nroots)] if spec.kind == N.KIND_DAVIDSON else None,
        "seconds_solve_device": res.device_ms_solve * 1e-3, "seconds_wall_with_setup": wall,
        "iterations_per_s": res.iterations / (res.device_ms_solve * 1e-3) if res.device_ms_solve else None,
        "handler_gbs_per_gpu": res.handler_bytes / res.handler_device_seconds / 1e9 if res.handler_device_seconds else None,
        "handler_device_seconds": res.handler_device_seconds,
        "gemm_inner_gbs": res.bytes_gemm_inner / res.seconds_gemm_inner / 1e9 if res.seconds_gemm_inner else None,
        "gemm_outer_gbs": res.bytes_gemm_outer / res.seconds_gemm_outer / 1e9 if res.seconds_gemm_outer else None,
        "streaming_gbs": res.bytes_blas1 / res.seconds_blas1 / 1e9 if res.seconds_blas1 else None,
        "residual_gbs": res.bytes_residual / res.seconds_residual / 1e9 if res.seconds_residual else None,
        "seconds_by_family": {"gemm_inner": res.seconds_gemm_inner, "gemm_outer": res.seconds_gemm_outer,
                              "blas1": res.seconds_blas1, "residual": res.seconds_residual},
        "seconds_action": res.seconds_action,
        "peak_vectors": peak / (8.0 * max(nloc, 1)), "peak_gb_per_gpu": peak / 1e9,
        "launches": int(res.kernel_launches),
        "calls": {"dot": int(res.n_dot), "gemm_inner": int(res.n_gemm_inner), "gemm_outer": int(res.n_gemm_outer),
                  "axpy": int(res.n_axpy), "scal": int(res.n_scal), "copy": int(res.n_copy), "fill": int(res.n_fill)},
    }
    if rank == 0:
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
