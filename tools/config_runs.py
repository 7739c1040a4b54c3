"""BASELINE.json's configurations on the CUDA backend at their stated sizes, each gated by what is checkable at size.

Used by tools/run_config.py (command line) and by bench.py (the `configs` key of the bench line). Everything here is
harness and checker code: the solves run through the C ABI of include/itsolv_b200_harness.h, the checks are independent
of it (a torch restatement of the operator, the known solutions the inputs were built from, and the reference's own CPU
path at a size the host finishes in seconds).

Gates of a run (`gates` in the record, `gated` = all of them hold):
  converged            the solver's own flag, errors below the threshold
  iterations_match     same iteration count as the reference's std::vector path on the same configuration at n_small rows
                       (the low end of the spectrum and the scaled inputs depend on the first rows only)
  eigenvalues_match    Davidson: eigenvalues within 1e-10 relative of that reference run
  residual_ok          ||A x - lambda x|| / ||x||  (Davidson), ||A x - b|| / ||b||  (LinearEquations), ||A (v - t)||  (DIIS)
                       from an operator application written in torch, not the harness kernels: <= 1e-7
  solution_ok          LinearEquations / DIIS: max |x - x_known| <= 1e-6 max(1, |x_known|)
"""
from __future__ import annotations

import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

B, EPS = 4, 1e-3

# name -> harness spec at the size BASELINE.json states
CONFIGS = {
    "c2": dict(n=10_000_000, kind="davidson", nroots=4, hermitian=1),
    "c3": dict(n=200_000_000, kind="lineq", nroots=8, hermitian=1, max_size_qspace=24),
    "c4": dict(n=2_000_000_000, kind="davidson", nroots=16, hermitian=1, max_size_qspace=8, nbuffers=8),
    "c5a": dict(n=500_000_000, kind="davidson", nroots=4, hermitian=1, max_p=500),
    "c5b": dict(n=500_000_000, kind="diis", max_size_qspace=8),
}
# vectors resident at the high-water mark of the fused path, measured at small n (tools/run_config.py --memory-table);
# a configuration is feasible on N GPUs when peak_vectors * 8 n / N fits the memory the driver reports free
N_SMALL_DEFAULT = 1_000_000


def _kind(name):
    from iterative_solver_b200 import _native as N
    return {"davidson": N.KIND_DAVIDSON, "lineq": N.KIND_LINEQ, "diis": N.KIND_DIIS}[name]


def make_spec(kw, **extra):
    from iterative_solver_b200 import harness as H
    kw = dict(kw)
    kw.update(extra)
    kw["kind"] = _kind(kw["kind"]) if isinstance(kw["kind"], str) else kw["kind"]
    return H.make_spec(half_bandwidth=B, eps=EPS, **kw)


# ---- independent checker: the banded operator written in torch (chunked, rows [start, start + nloc) of a global vector)

def _edges(x, b, rank, world):
    """b rows from each neighbouring shard (zeros at the ends of the global vector)"""
    import torch
    import torch.distributed as dist
    lo = torch.zeros(b, dtype=x.dtype, device=x.device)
    hi = torch.zeros(b, dtype=x.dtype, device=x.device)
    if world > 1:
        mine = torch.cat([x[:b], x[-b:]]).contiguous()
        allx = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allx, mine)
        if rank > 0:
            lo = allx[rank - 1][b:].clone()
        if rank < world - 1:
            hi = allx[rank + 1][:b].clone()
    return lo, hi


def torch_banded_apply(x, n, start, rank, world, b=B, eps=EPS, chunk=1 << 24):
    """y = A x for this shard: A(i,i) = i+1, A(i,j) = eps (1 + ((i+j) mod 7)) for 0 < |i-j| <= b"""
    import torch
    nloc = x.numel()
    lo, hi = _edges(x, b, rank, world)
    y = torch.empty_like(x)
    for c0 in range(0, nloc, chunk):
        c1 = min(nloc, c0 + chunk)
        # x over [c0 - b, c1 + b) with the neighbours' rows where the chunk touches the ends of the shard
        left = x[c0 - b:c0] if c0 >= b else torch.cat([lo[c0:], x[0:c0]])
        right = x[c1:c1 + b] if c1 + b <= nloc else torch.cat([x[c1:nloc], hi[:c1 + b - nloc]])
        xe = torch.cat([left, x[c0:c1], right])
        i = torch.arange(start + c0, start + c1, dtype=torch.int64, device=x.device)
        acc = (i + 1).to(torch.float64) * xe[b:b + (c1 - c0)]
        for d in range(1, b + 1):
            up = eps * (1 + ((2 * i + d) % 7)).to(torch.float64) * xe[b + d:b + d + (c1 - c0)]
            up = torch.where(i + d < n, up, torch.zeros_like(up))
            dn = eps * (1 + ((2 * i - d) % 7)).to(torch.float64) * xe[b - d:b - d + (c1 - c0)]
            dn = torch.where(i - d >= 0, dn, torch.zeros_like(dn))
            acc = acc + up + dn
        y[c0:c1] = acc
    return y


def known_solution(kind, k, start, nloc, device, legacy=False):
    """x_k of the LinearEquations inputs / the target t of the DIIS residual (include/itsolv_b200_harness.h, ITSOLV_RHS_*)"""
    import torch
    i = torch.arange(start, start + nloc, dtype=torch.int64, device=device)
    if kind == "diis":
        return torch.ones(nloc, dtype=torch.float64, device=device) if legacy else 1.0 / (i + 1).to(torch.float64)
    u = ((i * (k + 2) + k) % (2 * k + 5)).to(torch.float64) / float(2 * k + 5) - 0.5
    return u if legacy else (float(k + 1) + u) / (i + 1).to(torch.float64)


def _allsum(vals, world):
    import torch
    import torch.distributed as dist
    t = torch.tensor(vals, dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t)
    return t.cpu().numpy()


def _allmax(vals, world):
    import torch
    import torch.distributed as dist
    t = torch.tensor(vals, dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.cpu().numpy()


class DeviceRows:
    """nroots x nloc doubles from the context's own pool (so that they count in its high-water mark and need no host
    memory); row(k) is a torch view of the k-th vector"""

    def __init__(self, ctx, nroots, nloc, address=0):
        self.ctx, self.nroots, self.nloc = ctx, nroots, nloc
        self.address = address or ctx.alloc(nroots * max(nloc, 1))

    def row(self, k):
        import torch

        class Raw:
            pass
        raw = Raw()
        raw.__cuda_array_interface__ = {"shape": (self.nloc,), "typestr": "<f8", "version": 3,
                                        "data": (self.address + 8 * k * self.nloc, False)}
        return torch.as_tensor(raw, device="cuda")

    def free(self):
        if self.address:
            self.ctx.free(self.address)
            self.address = 0


def verify_solutions(kind, n, start, rows, nroots, eigenvalues, rank, world, legacy=False, ctx=None, detail=None):
    """residuals (and distances to the known solutions) of the solution vectors, one root at a time on the GPU;
    rows(k) returns the k-th solution as a torch tensor on the device"""
    res_norm, sol_err = [], []
    for k in range(nroots):
        x = rows(k)
        nloc = x.numel()
        if kind == "davidson":
            y = torch_banded_apply(x, n, start, rank, world)
            r = y - eigenvalues[k] * x
            num, den = _allsum([float((r * r).sum()), float((x * x).sum())], world)
            res_norm.append(float(np.sqrt(num / den)))
            if detail is not None and ctx is not None and world == 1:
                # the same residual with the harness' own operator kernel, and where the two operators differ most
                import torch
                y2 = torch.empty_like(x)
                ctx.banded_apply(x, y2, n, 0, B, EPS)
                ctx.synchronize()
                r2 = y2 - eigenvalues[k] * x
                detail.setdefault("harness_operator_residual", []).append(float(r2.norm() / x.norm()))
                d = (y - y2).abs()
                imax = int(d.argmax())
                detail.setdefault("operators_differ", []).append([float(d.max()), imax, float(x[imax]), float(y[imax])])
                imax = int(r.abs().argmax())
                detail.setdefault("largest_residual_entry", []).append([imax, float(r[imax]), float(x[imax])])
                del y2, r2, d
            del y
        elif kind == "lineq":
            xk = known_solution(kind, k, start, nloc, x.device, legacy)
            bk = torch_banded_apply(xk, n, start, rank, world)
            r = torch_banded_apply(x, n, start, rank, world) - bk
            num, den = _allsum([float((r * r).sum()), float((bk * bk).sum())], world)
            res_norm.append(float(np.sqrt(num / den)))
            err, scale = _allmax([float((x - xk).abs().max()), float(xk.abs().max())], world)
            sol_err.append(float(err / max(1.0, scale)))
            del xk, bk
        else:
            t = known_solution(kind, 0, start, nloc, x.device, legacy)
            r = torch_banded_apply(x - t, n, start, rank, world)
            num, = _allsum([float((r * r).sum())], world)
            res_norm.append(float(np.sqrt(num)))
            err, = _allmax([float((x - t).abs().max())], world)
            sol_err.append(float(err))
            del t
        del x, r
    return res_norm, sol_err


def reference_small(kw, n_small):
    """the reference's own std::vector path (oracle/_ref) on the same configuration at n_small rows; None when the
    reference build is not on this box"""
    import itsolv_oracle_lib
    o = itsolv_oracle_lib.load()
    if o.ref is None:
        return None
    kw = dict(kw)
    kw["n"] = n_small
    kw.pop("fused", None)
    t0 = time.perf_counter()
    res, _ = o.ref.solve(make_spec(kw))
    nroots = 1 if kw["kind"] == "diis" else kw["nroots"]
    return {"n": n_small, "iterations": int(res.iterations), "converged": int(res.converged),
            "eigenvalues": [float(res.eigenvalues[i]) for i in range(nroots)] if kw["kind"] == "davidson" else None,
            "seconds": time.perf_counter() - t0}


def run(ctx, name, rank, world, n=None, overrides=None, fused=None, verify=True, n_small=N_SMALL_DEFAULT,
        warm=True, reference=True):
    """One configuration: cold solve (populates the memory pool), warm solve (timed, device events), checks."""
    import torch
    import iterative_solver_b200 as pkg
    from iterative_solver_b200 import _native as N
    from iterative_solver_b200 import harness as H

    kw = dict(CONFIGS[name])
    if n is not None:
        kw["n"] = int(n)
    kw.update(overrides or {})
    kind = kw["kind"]
    if fused is None:
        fused = 1
    kw["fused"] = fused
    legacy = kw.get("rhs_kind", N.RHS_SCALED) == N.RHS_LEGACY
    spec = make_spec(kw)
    borders = pkg.distribution(spec.n, world)
    start, nloc = int(borders[rank]), int(borders[rank + 1] - borders[rank])
    nroots = 1 if kind == "diis" else kw["nroots"]
    ctx.set_profiling(True)
    ctx.mem_usage(reset_peak=True)
    # the solution vectors stay on the GPU, in memory of the context's pool that the harness takes only after the solver
    # has finished (itsolv_harness_problem_solve_device): nothing of size n goes through the host, and the high-water mark
    # below is the solver's own or (vectors alive at the end + solutions), whichever is larger
    sol = None
    problem = H.Problem(ctx, spec)
    t0 = time.perf_counter()
    res = problem.solve(spec)
    cold_s = time.perf_counter() - t0
    if verify:
        res, address = problem.solve_device(spec)
        sol = DeviceRows(ctx, nroots, nloc, address)
    elif warm:
        res = problem.solve(spec)
    problem.close()
    live, peak = ctx.mem_usage()
    ctx.set_profiling(False)
    eig = [float(res.eigenvalues[i]) for i in range(nroots)] if kind == "davidson" else None
    errors = [float(res.errors[i]) for i in range(nroots)]
    secs = res.device_ms_solve * 1e-3
    rec = {
        "config": name, "spec": {k: v for k, v in kw.items()}, "n_gpus": world, "n_local": nloc,
        "converged": int(res.converged), "iterations": int(res.iterations), "errors_max": max(errors),
        "eigenvalues": eig, "seconds_solve_device": secs, "seconds_first_solve_wall": cold_s,
        "iterations_per_s": res.iterations / secs if secs else None,
        "handler_gbs_per_gpu": res.handler_bytes / res.handler_device_seconds / 1e9 if res.handler_device_seconds else None,
        "handler_device_seconds": res.handler_device_seconds, "seconds_action": res.seconds_action,
        "family_gbs": {
            "gemm_inner": res.bytes_gemm_inner / res.seconds_gemm_inner / 1e9 if res.seconds_gemm_inner else None,
            "gemm_outer": res.bytes_gemm_outer / res.seconds_gemm_outer / 1e9 if res.seconds_gemm_outer else None,
            "blas1": res.bytes_blas1 / res.seconds_blas1 / 1e9 if res.seconds_blas1 else None,
            "residual": res.bytes_residual / res.seconds_residual / 1e9 if res.seconds_residual else None},
        "family_seconds": {"gemm_inner": res.seconds_gemm_inner, "gemm_outer": res.seconds_gemm_outer,
                           "blas1": res.seconds_blas1, "residual": res.seconds_residual},
        "peak_vectors": peak / (8.0 * max(nloc, 1)), "peak_gb_per_gpu": peak / 1e9, "launches": int(res.kernel_launches),
    }
    gates = {"converged": bool(res.converged)}
    if verify:
        ctx.mem_trim()  # the pool keeps every freed vector; the checker's torch temporaries need some of it back
        detail = {} if os.environ.get("ITSOLV_VERIFY_DETAIL") else None
        res_norm, sol_err = verify_solutions(kind, spec.n, start, sol.row, nroots, eig, rank, world, legacy, ctx, detail)
        rec["independent_residual_max"] = max(res_norm)
        rec["independent_residuals"] = res_norm
        if detail:
            rec["verify_detail"] = detail
        gates["residual_ok"] = max(res_norm) <= 1e-7
        if sol_err:
            rec["solution_error_max"] = max(sol_err)
            gates["solution_ok"] = max(sol_err) <= 1e-6
        sol.free()
        torch.cuda.empty_cache()
        ctx.mem_trim()
    if reference and n_small:
        small = None
        if rank == 0:
            small = reference_small(kw, min(n_small, spec.n))
        if world > 1:
            import torch.distributed as dist
            box = [small]
            dist.broadcast_object_list(box, src=0)
            small = box[0]
        if small is not None:
            rec["reference_small"] = small
            gates["iterations_match"] = small["iterations"] == int(res.iterations) and small["converged"] == int(res.converged)
            if eig:
                dev = max(abs(a / b - 1) for a, b in zip(eig, small["eigenvalues"]))
                rec["eigenvalues_vs_reference_small"] = dev
                gates["eigenvalues_match"] = dev <= 1e-10
    rec["gates"] = gates
    rec["gated"] = all(gates.values())
    return rec
