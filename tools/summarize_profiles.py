#!/usr/bin/env python
"""Turns the raw ncu CSVs / reports brought back in gpurun_out/ into the small tracked summaries under profiles/."""
import collections
import csv
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
SRC = os.path.join(ROOT, "gpurun_out")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r01"


# kernel families as bench.py accounts them (itsolv_counters): name prefixes of the kernels of each
FAMILIES = {
    "gemm_inner": ("gemm_inner",),
    "gemm_outer": ("gemm_outer",),
    "residual": ("davidson_residual",),
    "blas1": ("axpy", "scal", "copy", "fill", "precondition", "mgs_step", "shift", "elementwise"),
    "select": ("select_",),
    "harness_spmv": ("csr_apply", "banded_apply"),
}


def load(path):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    return list(csv.DictReader(lines))


def short(name):
    name = re.sub(r"^void ", "", name)
    name = name.replace("itsolv::", "")
    return re.sub(r"\(.*", "", name)


def to_us(v, unit):
    v = float(v.replace(",", ""))
    return {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}.get(unit, v)


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def per_launch(rows):
    by = collections.OrderedDict()
    for r in rows:
        d = by.setdefault(r["ID"], {"name": short(r["Kernel Name"]), "grid": r["Grid Size"], "block": r["Block Size"]})
        m = r["Metric Name"]
        if m == "gpu__time_duration.sum":
            d["us"] = to_us(r["Metric Value"], r["Metric Unit"])
        elif m.startswith("dram__bytes"):
            d[m] = to_bytes(r["Metric Value"], r["Metric Unit"])
    return [d for d in by.values() if "at::" not in d["name"] and "us" in d]


def table(launches, path, with_dram=False):
    agg = collections.OrderedDict()
    for d in launches:
        a = agg.setdefault((d["name"], d["grid"], d["block"]), {"n": 0, "us": 0.0, "min": 1e30, "rd": 0.0, "wr": 0.0})
        a["n"] += 1
        a["us"] += d["us"]
        a["min"] = min(a["min"], d["us"])
        a["rd"] += d.get("dram__bytes_read.sum", 0.0)
        a["wr"] += d.get("dram__bytes_write.sum", 0.0)
    total = sum(a["us"] for a in agg.values())
    with open(path, "w", newline="") as f:
        w = csv.writer(f)
        hdr = ["kernel", "grid", "block", "launches", "total_us", "avg_us", "min_us", "share_of_kernel_time"]
        if with_dram:
            hdr += ["avg_dram_read_MB", "avg_dram_write_MB", "dram_GBps"]
        w.writerow(hdr)
        for (name, grid, block), a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
            row = [name, grid, block, a["n"], f"{a['us']:.1f}", f"{a['us']/a['n']:.2f}", f"{a['min']:.2f}",
                   f"{a['us']/total:.4f}"]
            if with_dram:
                row += [f"{a['rd']/a['n']/1e6:.2f}", f"{a['wr']/a['n']/1e6:.2f}", f"{(a['rd']+a['wr'])/a['us']/1e3:.0f}"]
            w.writerow(row)
    return agg, total


def fresh(path, hours=6.0):
    """gpurun_out/ keeps what earlier calls (and earlier rounds) brought back: only files of the last few hours count"""
    import time
    return os.path.exists(path) and time.time() - os.path.getmtime(path) < hours * 3600


def main():
    os.makedirs(OUT, exist_ok=True)
    p = os.path.join(SRC, "opbench_launches.csv")
    if fresh(p):
        table(per_launch(load(p)), os.path.join(OUT, f"ncu_launches_opbench_{TAG}.csv"))
    p = os.path.join(SRC, "bench_dram.csv")
    if fresh(p):
        launches = per_launch(load(p))
        ours = launches[len(launches) // 2:]  # the second (timed) solve of `bench.py --steps 1 --warmup 1`
        agg, total = table(ours, os.path.join(OUT, f"ncu_dram_bench_{TAG}.csv"), with_dram=True)
        fam = {"n": 0, "us": 0.0, "bytes": 0.0}
        families = {k: {"launches": 0, "us": 0.0, "bytes": 0.0} for k in FAMILIES}
        for (name, _, _), a in agg.items():
            if name.startswith("gemm_inner"):
                fam["n"] += a["n"]
                fam["us"] += a["us"]
                fam["bytes"] += a["rd"] + a["wr"]
            for k, prefixes in FAMILIES.items():
                if name.startswith(prefixes):
                    families[k]["launches"] += a["n"]
                    families[k]["us"] += a["us"]
                    families[k]["bytes"] += a["rd"] + a["wr"]
        for k, v in families.items():
            v["share_of_kernel_time"] = v["us"] / total
            v["dram_bytes_per_launch"] = v["bytes"] / max(v["launches"], 1)
            v["avg_us"] = v["us"] / max(v["launches"], 1)
            v["dram_GBps"] = v["bytes"] / v["us"] / 1e3 if v["us"] else 0.0
        with open(os.path.join(OUT, f"ncu_dram_bench_{TAG}.json"), "w") as f:
            json.dump({"command": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
                                  "--clock-control none python bench.py --steps 1 --warmup 1 --min-warmup 1 --no-e2e "
                                  "--no-cpu-baseline (second solve)",
                       "kernel_time_us_one_solve": total, "launches_one_solve": sum(a["n"] for a in agg.values()),
                       "families": families,
                       "gemm_inner_family": {"launches": fam["n"], "share_of_kernel_time": fam["us"] / total,
                                             "dram_bytes_per_launch": fam["bytes"] / max(fam["n"], 1),
                                             "avg_us": fam["us"] / max(fam["n"], 1)}}, f, indent=1)
    rep = os.path.join(SRC, "prof_gemm_inner.ncu-rep")
    if fresh(rep):
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr, units = rows[0], rows[1]
        want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum",
                "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
                "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
                "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
                "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
                "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
                "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
                "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active"]
        idx = [hdr.index(w) for w in want if w in hdr]
        with open(os.path.join(OUT, f"ncu_full_gemm_inner_{TAG}.csv"), "w", newline="") as f:
            w = csv.writer(f)
            w.writerow([hdr[i] + (f" [{units[i]}]" if units[i] else "") for i in idx])
            for r in rows[2:]:
                w.writerow([short(r[i]) if hdr[i] == "Kernel Name" else r[i] for i in idx])
    # raw page of a `ncu --set full` capture exported on the GPU box (`ncu -i … --page raw --csv`): one row per launch
    for src_name, out_name in (("bench_full_raw.csv", f"ncu_full_bench_{TAG}.csv"),
                               ("fused_raw.csv", f"ncu_full_fused_{TAG}.csv")):
        p = os.path.join(SRC, src_name)
        if not fresh(p):
            continue
        rows = list(csv.reader(open(p)))
        rows = [r for r in rows if len(r) > 10 and (r[0] == "ID" or r[0] == "" or r[0].isdigit())]
        hdr, units, data = rows[0], rows[1], rows[2:]
        want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum",
                "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
                "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
                "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
                "smsp__issue_active.avg.pct_of_peak_sustained_active",
                "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
                "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
                "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
                "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
                "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
                "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
                "smsp__inst_executed.sum"]
        idx = [hdr.index(w) for w in want if w in hdr]
        with open(os.path.join(OUT, out_name), "w", newline="") as f:
            w = csv.writer(f)
            w.writerow([hdr[i] + (f" [{units[i]}]" if units[i] else "") for i in idx])
            for r in data:
                w.writerow([short(r[i]) if hdr[i] == "Kernel Name" else r[i] for i in idx])
    p = os.path.join(SRC, f"opbench_{TAG}.json")
    if fresh(p):
        shutil.copy(p, os.path.join(OUT, f"opbench_{TAG}.json"))
    p = os.path.join(SRC, "bench.log")
    if fresh(p):
        lines = [l for l in open(p) if l.startswith("{")]
        if lines:
            with open(os.path.join(OUT, f"bench_{TAG}.json"), "w") as f:
                json.dump(json.loads(lines[-1]), f, indent=1)
    print(os.listdir(OUT))


if __name__ == "__main__":
    main()
