#!/usr/bin/env python
"""Benchmark of the hot path named by BASELINE.json: Davidson iterations/s and subspace-update HBM GB/s.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference --steps K --warmup W    (the reference's own CPU path on the host cores)

Workload at N=1 = BASELINE.json configs[1]: synthetic diagonally dominant banded symmetric CSR matrix, n = 1e7, 4 roots,
LinearEigensystemDavidson FP64, hermitian, R/Q containers DistrArrayCUDA, handlers ArrayHandlerCUDA. At N>1 every GPU
keeps a 1e7-row shard (weak scaling; vectors row-sharded, partial Gram matrices all-reduced over NCCL).
A step is one complete solve() (operator resident in HBM). Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

N_PER_GPU = 10_000_000
NROOTS = 4
HALF_BANDWIDTH = 4
EPS = 1e-3
NOMINAL_HBM_GBS = 8000.0  # the north star's "B200's ~8 TB/s"; the measured copy rate is MEASURED_PEAKS.json
METRIC = "davidson_iterations_per_s"
UNIT = "iterations/s"


def workload_name(n_per_gpu, nroots):
    return (f"BASELINE.json configs[1]: banded symmetric CSR operator (half bandwidth {HALF_BANDWIDTH}, eps {EPS}), "
            f"n={n_per_gpu:.0e} rows per GPU, {nroots} roots, LinearEigensystemDavidson FP64, hermitian")


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(family):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the roofline kernel family, from the newest committed
    ncu capture of this same command (profiles/ncu_dram_bench_rNN.json, written by tools/summarize_profiles.py)"""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "ncu_dram_bench_r*.json")))
    if not files:
        return None, None
    with open(files[-1]) as f:
        d = json.load(f)
    fam = d.get("families", {}).get(family)
    if fam is None or not fam.get("launches"):
        return None, None
    return fam["dram_bytes_per_launch"], os.path.relpath(files[-1], ROOT)


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML in-process every few milliseconds (the timed
    region of the default run is a fraction of a second), nvidia-smi as the fallback"""
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index: int):
        self.index = index
        self.samples = []  # (sm_mhz, sm_max_mhz, [reason flags])
        self.stop_flag = threading.Event()
        self.thread = threading.Thread(target=self.run, daemon=True)
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if visible:
                ids = [v.strip() for v in visible.split(",") if v.strip()]
                if index < len(ids) and ids[index].isdigit():
                    phys = int(ids[index])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            self.source = "nvml"
        except Exception:
            self.source = "nvidia-smi"

    def sample_nvml(self):
        n = self.nvml
        sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
        get = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        mask = int(get(self.handle))
        bits = [n.nvmlClocksThrottleReasonHwSlowdown, n.nvmlClocksThrottleReasonHwThermalSlowdown,
                n.nvmlClocksThrottleReasonSwThermalSlowdown, n.nvmlClocksThrottleReasonSwPowerCap]
        self.samples.append((sm, self.sm_max, [bool(mask & b) for b in bits]))

    def sample_smi(self):
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                              "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
        parts = [p.strip() for p in out.strip().split(",")]
        if len(parts) >= 6:
            self.samples.append((float(parts[0]), float(parts[1]), [p.lower().startswith("active") for p in parts[2:6]]))

    def run(self):
        while not self.stop_flag.is_set():
            try:
                if self.nvml:
                    self.sample_nvml()
                else:
                    self.sample_smi()
            except Exception:
                pass
            self.stop_flag.wait(0.004 if self.nvml else 0.2)

    def sample_now(self):
        """one sample from the calling thread (the edges of the timed region: an NVML query takes tens of milliseconds
        while kernels are being launched, so the background thread gets few samples out of a 0.2 s region)"""
        try:
            if self.nvml:
                self.sample_nvml()
            else:
                self.sample_smi()
        except Exception:
            pass

    def __enter__(self):
        self.sample_now()
        self.thread.start()
        return self

    def __exit__(self, *a):
        self.stop_flag.set()
        self.thread.join(timeout=6)
        self.sample_now()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock samples (nvml and nvidia-smi unavailable)"]}
        sm = sorted(s[0] for s in self.samples)
        reasons = [n for i, n in enumerate(self.NAMES) if any(s[2][i] for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.samples[0][1], "reasons": reasons, "samples": len(sm),
                "source": self.source}


def reference_arm(args):
    """The reference's own CPU path (std::vector + ArrayHandlerIterable driven by its LinearEigensystemDavidson
    template, compiled in place into oracle/_ref) on the host cores. The path is single-threaded by construction
    (no OpenMP / threads anywhere in it), so cores = 1."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import itsolv_oracle_lib
    from iterative_solver_b200 import _native as N
    from iterative_solver_b200 import harness as H
    o = itsolv_oracle_lib.load()
    if o.ref is None:
        raise SystemExit("oracle/_ref/libitsolv_ref.so is missing: run __graft_entry__.build() where /root/reference exists")
    # Every step is one complete solve of the full workload (n = 1e7 rows, 4 roots), as many steps and warm-up steps as
    # asked for (one warm-up solve at most): ~14 s each on this class of host, so the driver's run (20 steps) ends in
    # about five minutes.
    # At N > 1 the GPU arm is weak-scaled (N shards of 1e7 rows, value = N x iterations / time); the path is linear in n
    # and single-threaded, so on N x 1e7 rows the reference makes 1/N of the iterations per second and its value in the
    # same unit is the one measured here.
    n_sample = args.n
    spec = H.make_spec(n_sample, kind=N.KIND_DAVIDSON, nroots=args.roots, hermitian=1, half_bandwidth=HALF_BANDWIDTH,
                       eps=EPS)
    # one warm-up solve at most: the path is deterministic, single-threaded CPU code (nothing to compile or cache beyond the
    # first touch of its vectors), and every solve costs ~14 s of the driver's time
    warmup = min(args.warmup, 1)
    for _ in range(warmup):
        o.ref.solve(spec)
    iterations, seconds = 0, 0.0
    for _ in range(args.steps):
        t0 = time.perf_counter()
        res, _ = o.ref.solve(spec)
        seconds += time.perf_counter() - t0
        iterations += res.iterations
    scale = 1.0
    value = iterations / seconds
    sample = (f"{args.steps} complete solve(s) of the full workload (n={n_sample} rows, {args.roots} roots) after "
              f"{warmup} warm-up solve(s) (of {args.warmup} asked for); single-threaded path, host has "
              f"{os.cpu_count()} logical cores")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": warmup, "ms_per_step": seconds / args.steps * 1e3 / scale, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.n, args.roots), "n_per_gpu": args.n, "nroots": args.roots,
                   "n_global": args.n * args.gpus, "sharding": f"rows/{args.gpus}" if args.gpus > 1 else "none",
                   "value_unit_note": "iterations/s x number of 1e7-row shards (weak scaling)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "iterations_per_solve": iterations / args.steps,
    }
    print(json.dumps(line), flush=True)


CPU_OPS = [("dot", 0, 1, 1, lambda n, k, m: 16 * n), ("axpy", 1, 1, 1, lambda n, k, m: 24 * n),
           ("scal", 2, 1, 1, lambda n, k, m: 16 * n), ("gemm_inner[4x4]", 3, 4, 4, lambda n, k, m: 8 * n * (k + m)),
           ("gemm_inner[4x16]", 3, 4, 16, lambda n, k, m: 8 * n * (k + m)),
           ("gemm_outer[4x4]", 4, 4, 4, lambda n, k, m: 8 * n * (k + 2 * m)),
           ("gemm_outer[16x4]", 4, 16, 4, lambda n, k, m: 8 * n * (k + 2 * m))]


def _cpu_op_worker(args):
    op, n, k, m, reps = args
    import itsolv_oracle_lib
    return itsolv_oracle_lib.load().ref.time_op(op, n, k, m, reps)


def cpu_ops(args):
    """Second CPU baseline (`--cpu-ops`): the reference's own handler operations (ArrayHandlerIterable over std::vector,
    oracle/_ref) per shape at the bench's n, on one core and on all host cores at once. The reference scales over cores
    with MPI ranks that own row blocks (DistrArrayMPI3, array/util/gemm.h:170-182); there is no MPI in this image, so the
    all-core line runs one process per core on its own row block (n / cores rows, at least 2e6) and leaves the k x m MPI_Allreduce out - an upper bound
    of what that path can reach on this host. Algorithmic bytes as for the GPU kernels."""
    import multiprocessing as mp
    cores = len(os.sched_getaffinity(0))
    n = args.n
    rows = []
    ctx = mp.get_context("fork")
    for name, op, k, m, nbytes in CPU_OPS:
        reps = 3
        one = _cpu_op_worker((op, n, k, m, reps))
        share = max(n // cores, 2_000_000)  # rows per process: vectors of >= 16 MB, beyond the private caches
        with ctx.Pool(cores) as pool:
            t_all = max(pool.map(_cpu_op_worker, [(op, share, k, m, reps)] * cores))
        rows.append({"op": name, "seconds_1_core": one, "gbs_1_core": nbytes(n, k, m) / one / 1e9,
                     "rows_per_process": share, "seconds_all_cores": t_all,
                     "gbs_all_cores": nbytes(share * cores, k, m) / t_all / 1e9})
    print(json.dumps({"impl": "reference", "what": "per-op CPU baseline of the handler contract", "n": n, "cores": cores,
                      "kind": "reference (ArrayHandlerIterable, oracle/_ref); all-core line: one process per core on "
                              "n/cores rows, no Allreduce", "ops": rows}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cpu-ops", action="store_true",
                    help="per-op timing of the reference's CPU handlers on 1 and on all host cores (no GPU work)")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=N_PER_GPU, help="rows per GPU")
    ap.add_argument("--roots", type=int, default=NROOTS)
    ap.add_argument("--path", default="fused", choices=["fused", "unfused"],
                    help="fused: LinearEigensystemDavidsonFused (batched passes, same algorithm and API); "
                         "unfused: the reference's LinearEigensystemDavidson class on the CUDA handlers, call for call")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-path", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the comparison with the reference at 1e6 rows per GPU")
    ap.add_argument("--no-configs", action="store_true",
                    help="skip BASELINE.json's other configurations at their stated sizes (the `configs` key)")
    ap.add_argument("--min-warmup", type=int, default=3, help="profiling runs only: allow fewer than 3 warm-up steps")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.cpu_ops:
        return cpu_ops(args)
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    import torch.distributed as dist

    import iterative_solver_b200 as pkg
    from iterative_solver_b200 import _native as N
    from iterative_solver_b200 import distributed as D
    from iterative_solver_b200 import harness as H

    rank, world, local_rank = D.env_rank_world()
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    if os.environ.get("ITSOLV_BENCH_BIND", "1") != "0":
        numa = D.bind_host_to_device(local_rank)  # before any pinned host buffer is allocated
    else:
        numa = {"bound": False, "why": "ITSOLV_BENCH_BIND=0"}
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = pkg.Context(local_rank)
    D.attach_communicator(ctx)

    n_global = args.n * world
    fused = 1 if args.path == "fused" else 0
    spec = H.make_spec(n_global, kind=N.KIND_DAVIDSON, nroots=args.roots, hermitian=1, half_bandwidth=HALF_BANDWIDTH,
                       eps=EPS, explicit_csr=1, fused=fused)
    borders = pkg.distribution(n_global, world)
    lo, hi = int(borders[rank]), int(borders[rank + 1])

    # ---- parity with the reference, outside the timed region: the same solver path on 1e6 rows per GPU (sharded over all
    # ranks) against the reference's own std::vector path on the same GLOBAL problem, run on rank 0's host cores
    parity = None
    if not args.no_parity:
        parity = parity_with_reference(ctx, rank, world, args.roots, fused)
        ctx.mem_trim()  # the timed problem starts from an empty pool, as it does without the pre-check

    # ---- device-resident leg: operator (stored CSR) in HBM before the timed region starts
    problem = H.Problem(ctx, spec)
    ctx.set_profiling(True)  # before the warm-up: the pool of CUDA events is built there, not in the first timed step
    for _ in range(max(args.warmup, args.min_warmup)):
        problem.solve(spec)
    lib = N.kernels()
    iterations, launches = 0, 0
    agg = dict(bytes=0.0, secs=0.0, bgi=0.0, sgi=0.0, bgo=0.0, sgo=0.0, bb1=0.0, sb1=0.0, brs=0.0, srs=0.0, action=0.0,
               cgi=0, cgo=0, cb1=0, crs=0)
    with ClockSampler(local_rank) as clocks:
        # the ranks leave this barrier together; starting the sampler thread (NVML) before it keeps its start-up time,
        # which differs from rank to rank, out of the first timed step
        lib.itsolv_comm_barrier(ctx.handle)
        torch.cuda.synchronize()
        ctx.timer_start()
        detail = []
        for _ in range(args.steps):
            t_step = time.perf_counter()
            res = problem.solve(spec)
            detail.append({"wall_ms": (time.perf_counter() - t_step) * 1e3, "device_ms": res.device_ms_solve,
                           "action_ms": res.seconds_action * 1e3, "handler_ms": res.handler_device_seconds * 1e3,
                           "gemm_inner_ms": res.seconds_gemm_inner * 1e3, "gemm_outer_ms": res.seconds_gemm_outer * 1e3,
                           "blas1_ms": res.seconds_blas1 * 1e3, "residual_ms": res.seconds_residual * 1e3})
            iterations += res.iterations
            launches += res.kernel_launches
            agg["bytes"] += res.handler_bytes
            agg["secs"] += res.handler_device_seconds
            agg["bgi"] += res.bytes_gemm_inner
            agg["sgi"] += res.seconds_gemm_inner
            agg["bgo"] += res.bytes_gemm_outer
            agg["sgo"] += res.seconds_gemm_outer
            agg["bb1"] += res.bytes_blas1
            agg["sb1"] += res.seconds_blas1
            agg["brs"] += res.bytes_residual
            agg["srs"] += res.seconds_residual
            agg["cgi"] += res.calls_gemm_inner
            agg["cgo"] += res.calls_gemm_outer
            agg["cb1"] += res.calls_blas1
            agg["crs"] += res.calls_residual
        ms = ctx.timer_stop()
        lib.itsolv_comm_barrier(ctx.handle)
        torch.cuda.synchronize()
    ctx.set_profiling(False)
    if os.environ.get("ITSOLV_BENCH_RANK_DETAIL"):  # per-rank, per-step record (where a multi-GPU step spends its time)
        with open(f"{os.environ['ITSOLV_BENCH_RANK_DETAIL']}.rank{rank}.json", "w") as f:
            json.dump({"rank": rank, "world": world, "host_binding": numa, "steps": detail}, f)
    converged, eig = res.converged, [res.eigenvalues[i] for i in range(args.roots)]
    # the other driver path on the same operator, for the record (same timing protocol, fewer steps)
    other = None
    if not args.no_other_path:
        spec_o = H.make_spec(n_global, kind=N.KIND_DAVIDSON, nroots=args.roots, hermitian=1,
                             half_bandwidth=HALF_BANDWIDTH, eps=EPS, explicit_csr=1, fused=1 - fused)
        for _ in range(2):
            problem.solve(spec_o)
        lib.itsolv_comm_barrier(ctx.handle)
        o_steps, o_iter, o_launch = max(1, min(args.steps, 3)), 0, 0
        ctx.timer_start()
        for _ in range(o_steps):
            r_o = problem.solve(spec_o)
            o_iter += r_o.iterations
            o_launch += r_o.kernel_launches
        o_ms = float(ctx.allreduce_host(np.array([ctx.timer_stop()]), op_max=True)[0])
        other = {"path": "unfused" if fused else "fused", "value": world * o_iter / (o_ms * 1e-3), "unit": UNIT,
                 "ms_per_step": o_ms / o_steps, "gpu_launches_per_step": o_launch / o_steps,
                 "iterations_per_solve": o_iter / o_steps,
                 "eigenvalues_agree_to": max(abs(r_o.eigenvalues[i] / eig[i] - 1) for i in range(args.roots))}
    ms_max = float(ctx.allreduce_host(np.array([ms]), op_max=True)[0])
    sums = ctx.allreduce_host(np.array([agg["bytes"], agg["bgi"], agg["bgo"], agg["bb1"], float(launches), agg["brs"],
                                        float(agg["cgi"]), float(agg["cgo"]), float(agg["cb1"]), float(agg["crs"])]))
    maxs = ctx.allreduce_host(np.array([agg["secs"], agg["sgi"], agg["sgo"], agg["sb1"], agg["srs"]]), op_max=True)
    value = world * iterations / (ms_max * 1e-3)  # shard-iterations per second: every rank iterates over its 1e7-row shard
    problem.close()

    # ---- end-to-end leg: the caller owns the operator as host CSR (pinned); every step uploads it, solves, and downloads
    # eigenvalues and the solution vectors
    e2e = None
    if not args.no_e2e:
        row_ptr, col, val, diag = H.banded_csr_host(n_global, HALF_BANDWIDTH, EPS, lo, hi)
        pinned = [torch.from_numpy(a).pin_memory() for a in (row_ptr, col, val, diag)]
        row_ptr, col, val, diag = [t.numpy() for t in pinned]
        spec_e = H.make_spec(n_global, kind=N.KIND_DAVIDSON, nroots=args.roots, hermitian=1,
                             half_bandwidth=HALF_BANDWIDTH, eps=EPS, fused=fused)
        sol = torch.empty((args.roots, hi - lo), dtype=torch.float64).pin_memory()
        sol_np = sol.numpy()
        e_steps = max(1, min(args.steps, 3))

        def one_e2e():
            p = H.Problem(ctx, spec_e, csr=(row_ptr, col, val, diag))
            r = p.solve(spec_e, solutions=sol_np)
            p.close()
            return r

        one_e2e()
        lib.itsolv_comm_barrier(ctx.handle)
        torch.cuda.synchronize()
        e_iter = 0
        ctx.timer_start()
        for _ in range(e_steps):
            e_iter += one_e2e().iterations
        e_ms = ctx.timer_stop()
        lib.itsolv_comm_barrier(ctx.handle)
        e_ms = float(ctx.allreduce_host(np.array([e_ms]), op_max=True)[0])
        h2d = row_ptr.nbytes + col.nbytes + val.nbytes + diag.nbytes
        d2h = sol_np.nbytes + 8 * args.roots * 2
        e2e = {"value": world * e_iter / (e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "steps": e_steps, "ms_per_step": e_ms / e_steps, "host_binding": numa,
               "h2d_gbs_per_gpu_implied": h2d / max(e_ms / e_steps * 1e-3 - ms_max / args.steps * 1e-3, 1e-9) / 1e9,
               "api": "itsolv_harness_problem_create(host CSR) + itsolv_harness_problem_solve(host solutions)"}
        del pinned, sol

    # ---- CPU baseline beside it (rank 0, single GPU runs only): the reference's own std::vector path, full workload
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import itsolv_oracle_lib
        o = itsolv_oracle_lib.load()
        if o.ref is not None:
            spec_c = H.make_spec(args.n, kind=N.KIND_DAVIDSON, nroots=args.roots, hermitian=1,
                                 half_bandwidth=HALF_BANDWIDTH, eps=EPS)
            t0 = time.perf_counter()
            rres, _ = o.ref.solve(spec_c)
            dt = time.perf_counter() - t0
            same = (rres.iterations * args.steps == iterations and
                    max(abs(rres.eigenvalues[i] / eig[i] - 1) for i in range(args.roots)) <= 1e-10)
            cpu = {"value": rres.iterations / dt, "unit": UNIT, "cores": 1, "kind": "reference",
                   "sample": f"1 solve of the full workload (n={args.n}, {args.roots} roots) by the reference's "
                             f"std::vector/ArrayHandlerIterable path, {dt:.1f} s; host has {os.cpu_count()} logical cores, "
                             f"the path is single-threaded", "parity_with_gpu_run": bool(same)}

    # ---- BASELINE.json's other configurations at their stated sizes, gated (tools/config_runs.py); which ones depends on N
    configs = None
    if not args.no_configs and os.environ.get("ITSOLV_BENCH_CONFIGS", "1") != "0":
        configs = run_configs(ctx, rank, world)

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        # kernel families of the subspace path (itsolv_counters): algorithmic bytes, event time and calls of each, summed
        # over the timed region; the roofline object describes the one with the largest share of the device time
        families = {
            "gemm_inner": dict(kernel="gemm_inner_kernel / gemm_inner_direct_kernel (Gram blocks, dot)",
                               bytes=sums[1], secs=maxs[1], calls=sums[6]),
            "gemm_outer": dict(kernel="gemm_outer_kernel (subspace -> full space expansion, projection)",
                               bytes=sums[2], secs=maxs[2], calls=sums[7]),
            "blas1": dict(kernel="streaming kernels (mgs_step_dots, axpy/scal/copy/fill, preconditioner)",
                          bytes=sums[3], secs=maxs[3], calls=sums[8]),
            "residual": dict(kernel="davidson_residual_kernel (solution + residual + norms + preconditioner in one pass)",
                             bytes=sums[5], secs=maxs[4], calls=sums[9]),
        }
        dominant = max(families, key=lambda k: families[k]["secs"])
        dom = families[dominant]
        dom_gbs = agg_gbs(dom["bytes"] / world, dom["secs"])
        traffic, traffic_src = ncu_traffic(dominant)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, args.min_warmup),
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.n, args.roots), "n_per_gpu": args.n, "n_global": n_global,
                       "nroots": args.roots, "sharding": f"rows/{world}" if world > 1 else "none",
                       "driver_path": ("fused (LinearEigensystemDavidsonFused: the reference's algorithm and API, O(n) work "
                                       "batched)" if fused else "unfused (the reference's LinearEigensystemDavidson class)"),
                       "l2": "inputs larger than L2 (each vector 80 MB, ~40 live vectors)",
                       "value_unit_note": "iterations/s x number of 1e7-row shards (weak scaling)"},
            "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": int(sums[4]),
            "roofline": {"bound": "hbm", "kernel": dom["kernel"], "family": dominant,
                         "achieved": dom_gbs, "peak": peak, "unit": "GB/s", "frac": dom_gbs / peak if peak else None,
                         "peak_source": peak_src, "peak_nominal": NOMINAL_HBM_GBS,
                         "frac_nominal": dom_gbs / NOMINAL_HBM_GBS, "traffic": traffic, "traffic_source": traffic_src,
                         "algorithmic_bytes_per_launch": dom["bytes"] / max(dom["calls"], 1.0),
                         "launch_share_of_handler_time": dom["secs"] / maxs[0] if maxs[0] else None,
                         "families": {k: {"gbs": agg_gbs(v["bytes"] / world, v["secs"]),
                                          "frac": agg_gbs(v["bytes"] / world, v["secs"]) / peak if peak else None,
                                          "share_of_handler_time": v["secs"] / maxs[0] if maxs[0] else None,
                                          "launches": int(v["calls"])} for k, v in families.items()}},
            "cpu_baseline": cpu,
            "subspace_update": {"gbs_per_gpu": agg_gbs(sums[0] / world, maxs[0]),
                                "frac_of_measured_hbm": agg_gbs(sums[0] / world, maxs[0]) / peak,
                                "frac_of_nominal_hbm": agg_gbs(sums[0] / world, maxs[0]) / NOMINAL_HBM_GBS,
                                "gemm_outer_gbs": agg_gbs(sums[2] / world, maxs[2]),
                                "streaming_gbs": agg_gbs(sums[3] / world, maxs[3]),
                                "device_seconds_per_step": maxs[0] / args.steps},
            "iterations_per_solve": iterations / args.steps, "converged": int(converged), "eigenvalues": eig,
            "driver_path": args.path, "other_driver_path": other,
            "parity_with_reference": parity, "configs": configs,
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def parity_with_reference(ctx, rank, world, nroots, fused, rows_per_gpu=1_000_000):
    """The bench's solver path on rows_per_gpu x world rows, sharded over all ranks, against the reference's CPU path on
    the same global problem (rank 0): iteration count, convergence, eigenvalues to 1e-10."""
    from iterative_solver_b200 import _native as N
    from iterative_solver_b200 import harness as H
    n_global = rows_per_gpu * world
    kw = dict(kind=N.KIND_DAVIDSON, nroots=nroots, hermitian=1, half_bandwidth=HALF_BANDWIDTH, eps=EPS)
    res, _ = H.solve(ctx, H.make_spec(n_global, fused=fused, explicit_csr=1, **kw))
    out = {"n_global": n_global, "rows_per_gpu": rows_per_gpu, "iterations": int(res.iterations),
           "converged": int(res.converged), "reference": None, "ok": None}
    if rank == 0:
        import itsolv_oracle_lib
        o = itsolv_oracle_lib.load()
        if o.ref is not None:
            t0 = time.perf_counter()
            rres, _ = o.ref.solve(H.make_spec(n_global, **kw))
            dev = max(abs(res.eigenvalues[i] / rres.eigenvalues[i] - 1) for i in range(nroots))
            out["reference"] = {"iterations": int(rres.iterations), "converged": int(rres.converged),
                                "seconds": time.perf_counter() - t0, "kind": "oracle/_ref (the reference's own templates)"}
            out["eigenvalues_max_rel_dev"] = dev
            out["ok"] = bool(rres.iterations == res.iterations and rres.converged == res.converged and dev <= 1e-10)
    return out


def config_plan(world):
    """(name, overrides) of the configurations run at `world` GPUs. C4's stated shape (16 roots, 8 buffers, Q capped at 8)
    needs 8 GPUs for n = 2e9; on 4 and 2 GPUs the largest (roots, buffers, Q cap) whose high-water mark fits is run instead
    (profiles/memory_table_r02.jsonl)."""
    if world == 1:
        return [("c3", {})]
    if world == 2:
        return [("c3", {}), ("c4", dict(nroots=2, nbuffers=2, max_size_qspace=2)), ("c5a", {}), ("c5b", {})]
    if world == 4:
        return [("c4", dict(nroots=4, nbuffers=4, max_size_qspace=4)), ("c5a", {}), ("c5b", {})]
    return [("c4", {}), ("c5a", {}), ("c5b", {})]


def run_configs(ctx, rank, world):
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import config_runs as CR
    import iterative_solver_b200 as pkg
    records = []
    for name, ov in config_plan(world):
        try:
            # dry run at 2e5 rows: the high-water mark in vectors does not depend on n
            dry = CR.run(ctx, name, rank, world, n=200_000, overrides=ov, verify=True, reference=False, warm=False)
            n_full = CR.CONFIGS[name]["n"]
            nloc = int(pkg.distribution(n_full, world)[1])
            need = dry["peak_vectors"] * 8.0 * nloc
            ctx.mem_trim()
            try:
                import torch
                torch.cuda.empty_cache()
            except Exception:
                pass
            free_min = float(ctx.allreduce_host(np.array([-float(ctx.mem_info()[0])]), op_max=True)[0]) * -1.0
            if need + 6e9 > free_min:
                records.append({"config": name, "overrides": ov, "skipped": "does not fit",
                                "peak_vectors": dry["peak_vectors"], "need_gb_per_gpu": need / 1e9,
                                "free_gb_per_gpu": free_min / 1e9})
                continue
            rec = CR.run(ctx, name, rank, world, overrides=ov)
            rec["overrides"] = ov
            rec["need_gb_per_gpu_predicted"] = need / 1e9
            records.append(rec)
        except Exception as e:  # a failed configuration must not take the bench line with it
            records.append({"config": name, "overrides": ov, "error": repr(e)[:300]})
    return records


def agg_gbs(nbytes, seconds):
    return float(nbytes / seconds / 1e9) if seconds else 0.0


if __name__ == "__main__":
    main()
