#!/usr/bin/env python
"""Runs BASELINE.json's configurations on the CUDA backend, gated (tools/config_runs.py), one JSON line each (rank 0).
    python tools/run_config.py --config c3 [--n 2e7]                             (1 GPU)
    torchrun --nproc-per-node 8 tools/run_config.py --config c4 c5a c5b          (row-sharded)
    python tools/run_config.py --memory-table                                    (peak vectors of Davidson configurations)
configs: c2 Davidson n=1e7 4 roots | c3 LinearEquations n=2e8 8 RHS | c4 Davidson n=2e9 16 roots |
         c5a P-space (500) Davidson n=5e8 4 roots | c5b DIIS n=5e8"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tools")]
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import config_runs as CR  # noqa: E402
import iterative_solver_b200 as pkg  # noqa: E402
from iterative_solver_b200 import distributed as D  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", nargs="*", default=[], choices=sorted(CR.CONFIGS))
    ap.add_argument("--n", type=float, default=0, help="global rows instead of the configuration's own")
    ap.add_argument("--qspace", type=int, default=-1)
    ap.add_argument("--buffers", type=int, default=0)
    ap.add_argument("--roots", type=int, default=0)
    ap.add_argument("--max-p", type=int, default=-1)
    ap.add_argument("--max-iter", type=int, default=0)
    ap.add_argument("--threshold", type=float, default=0.0)
    ap.add_argument("--legacy-inputs", action="store_true", help="ITSOLV_RHS_LEGACY right-hand sides / DIIS target")
    ap.add_argument("--fused", type=int, default=1, help="1 fused driver (default), 0 the reference's classes call for call")
    ap.add_argument("--no-verify", action="store_true")
    ap.add_argument("--n-small", type=int, default=CR.N_SMALL_DEFAULT)
    ap.add_argument("--memory-table", action="store_true")
    ap.add_argument("--out", default="", help="append the JSON lines to this file as well")
    args = ap.parse_args()
    rank, world, local = D.env_rank_world()
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = pkg.Context(local)
    D.attach_communicator(ctx)

    def emit(rec):
        if rank == 0:
            line = json.dumps(rec)
            print(line, flush=True)
            if args.out:
                with open(args.out, "a") as f:
                    f.write(line + "\n")

    if args.memory_table:
        # high-water mark in vectors of the fused Davidson path, independent of n: what decides which (roots, buffers,
        # Q-space cap) fits on how many GPUs for n = 2e9
        for roots, buffers, q in ((16, 16, 8), (16, 8, 8), (16, 4, 4), (16, 2, 2), (8, 8, 8), (8, 4, 4), (8, 2, 2),
                                  (4, 4, 8), (4, 4, 4), (4, 2, 2), (2, 2, 2), (2, 1, 2), (1, 1, 2)):
            rec = CR.run(ctx, "c4", rank, world, n=200_000, overrides=dict(nroots=roots, nbuffers=buffers, max_size_qspace=q),
                         verify=False, n_small=0, warm=False)
            emit({"memory_table": True, "nroots": roots, "nbuffers": buffers, "max_size_qspace": q,
                  "peak_vectors": rec["peak_vectors"], "iterations": rec["iterations"], "converged": rec["converged"]})
    for name in args.config:
        ov = {}
        if args.qspace >= 0:
            ov["max_size_qspace"] = args.qspace
        if args.buffers:
            ov["nbuffers"] = args.buffers
        if args.roots:
            ov["nroots"] = args.roots
        if args.max_p >= 0:
            ov["max_p"] = args.max_p
        if args.max_iter:
            ov["max_iter"] = args.max_iter
        if args.threshold:
            ov["convergence_threshold"] = args.threshold
        if args.legacy_inputs:
            from iterative_solver_b200 import _native as N
            ov["rhs_kind"] = N.RHS_LEGACY
        rec = CR.run(ctx, name, rank, world, n=int(args.n) if args.n else None, overrides=ov, fused=args.fused,
                     verify=not args.no_verify, n_small=args.n_small)
        emit(rec)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
