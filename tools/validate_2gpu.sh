#!/bin/bash
# two ranks: the multi-GPU tests (both all-reduce modes) and the bench line at N=2 with the configurations that fit
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_multi_gpu.py -x -q -m gpu > gpurun_out/val2_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/val2_pytest.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/val2_bench_n2.json 2> gpurun_out/val2_bench_n2.err; echo "bench2 rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/val2_bench_n2.json") if l.startswith("{")][-1])
print(round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],2), d["parity_with_reference"]["ok"])
for c in d.get("configs") or []:
    print(c["config"], c["spec"]["n"], c.get("overrides"), "it", c.get("iterations"), "s", c.get("seconds_solve_device"), "GB/s", c.get("handler_gbs_per_gpu"), "gated", c.get("gated"), c.get("error"))
PY
