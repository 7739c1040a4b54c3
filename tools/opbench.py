#!/usr/bin/env python
"""Per-shape device timing of the contract's kernels through the C ABI (CUDA events on the context's stream, inputs
larger than L2 rotated between repetitions). Prints one line per shape: algorithmic GB/s and fraction of the measured
HBM copy rate. Tuning knobs: --opt NAME=VALUE (see itsolv_ctx_set_option)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import iterative_solver_b200 as pkg  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=10_000_000)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--opt", action="append", default=[])
    ap.add_argument("--only", default="")
    ap.add_argument("--pool", type=int, default=0, help="vectors in the rotating input pool (default 160 at n <= 1e7)")
    ap.add_argument("--json", default="")
    ap.add_argument("--sweep", action="append", default=[], help="NAME=v1,v2,...: rerun the selected ops for each value")
    args = ap.parse_args()
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    ctx = pkg.Context(0)
    ctx.init_comm(0, 1, b"\0" * 128)
    for o in args.opt:
        k, v = o.split("=")
        ctx.set_option(k, int(v))
    n = args.n
    g = torch.Generator(device="cuda").manual_seed(0)
    nvec = args.pool if args.pool > 0 else (160 if n <= 10_000_000 else 40)
    pool = [torch.randn(n, dtype=torch.float64, device="cuda", generator=g) for _ in range(nvec)]
    torch.cuda.synchronize()
    cursor = [0]

    def take(k):
        out = []
        for _ in range(k):
            out.append(pool[cursor[0] % nvec])
            cursor[0] += 1
        return out

    results = []

    def run(name, bytes_per_call, fn):
        if args.only and not any(pat in name for pat in args.only.split(",")):
            return
        fn()  # warm-up (also sets function attributes)
        fn()
        fn()
        ctx.synchronize()
        ctx.timer_start()
        for _ in range(args.reps):
            fn()
        ms = ctx.timer_stop() / args.reps
        # the same calls again with an event pair around each call's kernels: kernel time without host gaps
        ctx.reset_counters()
        ctx.set_profiling(True)
        for _ in range(args.reps):
            fn()
        ctx.synchronize()
        kms = ctx.counters().device_seconds * 1e3 / args.reps
        ctx.set_profiling(False)
        gbs = bytes_per_call / ms / 1e6
        kgbs = bytes_per_call / kms / 1e6 if kms else 0.0
        results.append({"op": name, "us": ms * 1e3, "gbs": gbs, "frac": gbs / peak, "kernel_us": kms * 1e3,
                        "kernel_gbs": kgbs, "kernel_frac": kgbs / peak})
        print(f"{name:22s} call {ms*1e3:9.1f} us {gbs:7.0f} GB/s {gbs/peak:5.2f} | kernels {kms*1e3:9.1f} us "
              f"{kgbs:7.0f} GB/s {kgbs/peak:5.2f} of measured HBM copy", flush=True)

    lib = ctx.lib
    from iterative_solver_b200.api import _ptr_array, _dbl
    sweeps = [("", [None])]
    if args.sweep:
        sweeps = []
        for sw in args.sweep:
            k, vs = sw.split("=")
            sweeps.append((k, [int(v) for v in vs.split(",")]))
    for sweep_name, sweep_values in sweeps:
      for sweep_value in sweep_values:
        if sweep_name:
            ctx.set_option(sweep_name, sweep_value)
            print(f"--- {sweep_name}={sweep_value}", flush=True)
            results.append({"option": sweep_name, "value": sweep_value})
        run_all(ctx, lib, args, n, nvec, take, run, _ptr_array, _dbl)
        if sweep_name:
            ctx.set_option(sweep_name, 0)
    if args.json:
        with open(args.json, "w") as f:
            json.dump({"n": n, "options": args.opt, "peak_gbs": peak, "results": results}, f, indent=1)
    ctx.close()


def run_all(ctx, lib, args, n, nvec, take, run, _ptr_array, _dbl):
    out = np.zeros(128 * 128)

    def gi(k, m):
        def f():
            xs, ys = take(k), take(m)
            # asynchronous variant: launch only (no host sync inside the timed loop would need a C entry; use full call)
            ctx._check(lib.itsolv_gemm_inner_f64(ctx.handle, _ptr_array(xs), k, _ptr_array(ys), m, n, _dbl(out)))
        return f

    def go(k, m):
        alpha = np.random.default_rng(0).standard_normal((k, m)) * 1e-3

        def f():
            ctx.gemm_outer(alpha, take(k), take(m))
        return f

    run("spmv_generated[b=4]", 16 * n, lambda: (lambda v: ctx.banded_apply(v[0], v[1], n, 0, 4, 1e-3))(take(2)))
    # select (initial guess / P-space choice): values spread over many binary exponents (the diagonal i + 1 on rank 0), the
    # shard of rank 7 of 8 of the same diagonal (one exponent, common leading mantissa bits), random values
    if not args.only or "select" in args.only:
        wide = torch.arange(1, n + 1, dtype=torch.float64, device="cuda")
        shard = wide + 7.0 * n
        run("select[4, diagonal]", 3 * 8 * n, lambda: ctx.select(wide, 4))
        run("select[4, shard 7/8]", 3 * 8 * n, lambda: ctx.select(shard, 4, global_offset=7 * n))
        run("select[500, random]", 3 * 8 * n, lambda: ctx.select(take(1)[0], 500))
        del wide, shard
    run("fill", 8 * n, lambda: ctx.fill(1.0, take(1)[0]))
    run("scal", 16 * n, lambda: ctx.scal(1.0000001, take(1)[0]))
    run("copy", 16 * n, lambda: ctx.copy(*take(2)))
    run("axpy", 24 * n, lambda: ctx.axpy(1e-9, *take(2)))
    run("dot", 16 * n, lambda: ctx.dot(*take(2)))
    run("dot(x,x)", 8 * n, lambda: (lambda v: ctx.dot(v, v))(take(1)[0]))
    # denominators close to 1 so that the vectors, which are rewritten in place, keep their magnitude over the repetitions
    diag = 1.0 + 1e-3 * torch.rand(n, dtype=torch.float64, device="cuda")
    run("precondition[w=4]", 8 * n * 9, lambda: ctx.precondition(take(4), diag, [1e-4, 2e-4, 3e-4, 4e-4]))
    for k, m in [(4, 1), (1, 4), (4, 4), (4, 8), (4, 12), (4, 16), (4, 20), (8, 8), (16, 16), (16, 24), (16, 40), (16, 64),
                 (8, 100), (16, 128), (32, 32), (64, 64), (128, 128)]:
        if (k + m) > nvec:
            continue
        run(f"gemm_inner[{k}x{m}]", 8 * n * (k + m), gi(k, m))
    for k, m in [(1, 4), (4, 4), (8, 4), (12, 4), (20, 4), (24, 16), (40, 16), (100, 16), (16, 40)]:
        if (k + m) > nvec:
            continue
        run(f"gemm_outer[{k}x{m}]", 8 * n * (k + 2 * m), go(k, m))

    # ---- kernels of the fused driver path
    def ds(k, m):
        coef = np.random.default_rng(1).standard_normal((k, m)) * 1e-2
        lam = 1e-4 * np.arange(1, m + 1)
        outs = [torch.empty(n, dtype=torch.float64, device="cuda") for _ in range(m)]  # not fed back into the pool

        def f():
            ctx.davidson_residual(coef, take(k), take(k), lam, outs, diag=diag)
        return f

    for k, m in [(4, 4), (8, 4), (16, 4), (24, 8), (24, 16), (40, 16)]:
        if 2 * k + m > nvec:
            continue
        run(f"davidson_residual[{k}x{m}]", 8 * n * (2 * k + m + 1), ds(k, m))
    for m in (0, 1, 3, 7, 15):
        run(f"mgs_step_dots[1+{m}]", 16 * n * (m + 1),
            lambda m=m: (lambda v: ctx.mgs_step_dots(1.0000001, v[0], [1e-9] * m, v[1:]))(take(m + 1)))
    for k, m in [(4, 4), (12, 4), (24, 16)]:
        alpha = np.random.default_rng(2).standard_normal((k, m)) * 1e-3
        run(f"gemm_outer_scaled[{k}x{m}]", 8 * n * (k + 2 * m),
            lambda k=k, m=m, alpha=alpha: ctx.gemm_outer_scaled(alpha, take(k), take(m), [1.0000001] * m))
    # stored CSR operator, banded with 9 entries per row, all vectors of a working set in one pass over the matrix
    from iterative_solver_b200 import harness as H
    row_ptr, col, val, _ = H.banded_csr_host(n, 4, 1e-3)
    d_rp, d_col, d_val = torch.from_numpy(row_ptr).cuda(), torch.from_numpy(col).cuda(), torch.from_numpy(val).cuda()
    nnz = int(row_ptr[-1])
    for w in (1, 4, 8):
        run(f"spmv_csr[w={w}]", 12 * nnz + 8 * (n + 1) + 16 * n * w,
            lambda w=w: ctx.csr_apply_multi(d_rp.data_ptr(), d_col.data_ptr(), d_val, take(w), take(w), n, 0, 4))


if __name__ == "__main__":
    main()
