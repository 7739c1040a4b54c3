#!/bin/bash
# what the driver runs at round end, on one GPU: the GPU tests, smoke(), the default bench line
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/val1_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/val1_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/val1_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/val1_smoke.log | cut -c1-160
timeout 900 python bench.py > gpurun_out/val1_bench.json 2> gpurun_out/val1_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/val1_bench.json") if l.startswith("{")][-1])
fam=d["roofline"]["families"]; tot=d["subspace_update"]["device_seconds_per_step"]*1e3
print(round(d["ms_per_step"],3), "value", round(d["value"],1), "launches/step", d["gpu_launches"]/d["steps"], "e2e", round(d["e2e"]["ms_per_step"],2), "handler", round(tot,3), {k:(round(v["share_of_handler_time"]*tot,3), round(v["frac"],3)) for k,v in fam.items()})
print("roofline", d["roofline"]["family"], round(d["roofline"]["frac"],3), "parity", d["parity_with_reference"]["ok"], "cpu", d["cpu_baseline"]["value"], "configs", [(c["config"], c.get("gated")) for c in d["configs"]])
PY
