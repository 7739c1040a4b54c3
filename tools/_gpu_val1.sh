#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/val1_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/val1_pytest.log
timeout 300 python tools/opbench.py --only select --reps 10 > gpurun_out/val1_opbench_select.txt 2>&1; cat gpurun_out/val1_opbench_select.txt
timeout 900 python bench.py > gpurun_out/val1_bench.json 2> gpurun_out/val1_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/val1_bench.json",):
    d=json.loads([l for l in open(f) if l.startswith("{")][-1])
    fam=d["roofline"]["families"]; tot=d["subspace_update"]["device_seconds_per_step"]*1e3
    print(f, round(d["ms_per_step"],3), "launches/step", d["gpu_launches"]/d["steps"], "e2e", round(d["e2e"]["ms_per_step"],2), "handler", round(tot,3), {k:(round(v["share_of_handler_time"]*tot,3), round(v["frac"],3)) for k,v in fam.items()}, "other", round(tot*(1-sum(v["share_of_handler_time"] for v in fam.values())),3))
    print(d["parity_with_reference"]); print(d["cpu_baseline"])
    for c in d["configs"] or []: print(c["config"], c["spec"]["n"], c.get("iterations"), c.get("seconds_solve_device"), c.get("handler_gbs_per_gpu"), c.get("gated"))
PY
