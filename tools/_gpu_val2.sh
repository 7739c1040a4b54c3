#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/val2_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/val2_pytest.log
timeout 300 python tools/opbench.py --only select --reps 10 > gpurun_out/val2_opbench_select.txt 2>&1; cat gpurun_out/val2_opbench_select.txt
timeout 900 python bench.py --no-configs > gpurun_out/val2_bench_n1.json 2> gpurun_out/val2_bench_n1.err; echo "bench1 rc=$?"
ITSOLV_BENCH_RANK_DETAIL=gpurun_out/val2_detail timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 --no-configs > gpurun_out/val2_bench_n2.json 2> gpurun_out/val2_bench_n2.err; echo "bench2 rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/val2_bench_n1.json","gpurun_out/val2_bench_n2.json"):
    d=json.loads([l for l in open(f) if l.startswith("{")][-1])
    fam=d["roofline"]["families"]; tot=d["subspace_update"]["device_seconds_per_step"]*1e3
    print(f, round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],2), "handler", round(tot,3), {k:round(v["share_of_handler_time"]*tot,3) for k,v in fam.items()}, "other", round(tot*(1-sum(v["share_of_handler_time"] for v in fam.values())),3))
PY
