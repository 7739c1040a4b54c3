#!/bin/bash
# quick check after a change: GPU tests, smoke, one bench line with the per-family breakdown
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/val1_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/val1_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-configs > gpurun_out/val1_bench.json 2> gpurun_out/val1_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/val1_bench.json") if l.startswith("{")][-1])
fam=d["roofline"]["families"]; tot=d["subspace_update"]["device_seconds_per_step"]*1e3
print(round(d["ms_per_step"],3), "launches/step", d["gpu_launches"]/d["steps"], "e2e", round(d["e2e"]["ms_per_step"],2), "handler", round(tot,3), {k:(round(v["share_of_handler_time"]*tot,3), round(v["frac"],3)) for k,v in fam.items()}, "other", round(tot*(1-sum(v["share_of_handler_time"] for v in fam.values())),3))
print(d["other_driver_path"]["ms_per_step"], d["parity_with_reference"]["ok"])
PY
