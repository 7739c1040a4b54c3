#!/bin/bash
# one-GPU evidence of the round: tests, bench line, launch list with DRAM bytes (only after the plain run exited 0),
# per-shape timings, configurations that fit one GPU, per-op CPU baseline
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/final_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/final_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2> gpurun_out/final_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --steps 1 --warmup 1 --min-warmup 1 --no-e2e --no-cpu-baseline --no-parity --no-configs --no-other-path > gpurun_out/final_plain.log 2>&1 && \
  timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/bench_dram.csv \
    python bench.py --steps 1 --warmup 1 --min-warmup 1 --no-e2e --no-cpu-baseline --no-parity --no-configs --no-other-path > gpurun_out/final_ncu.log 2>&1; echo "ncu rc=$?"
timeout 900 python tools/opbench.py --json gpurun_out/opbench_r02.json > gpurun_out/opbench_r02.txt 2>&1; echo "opbench rc=$?"
rm -f gpurun_out/final_configs.jsonl
timeout 900 python tools/run_config.py --config c4 --n 2.5e8 --out gpurun_out/final_configs.jsonl > /dev/null 2> gpurun_out/final_configs.err; echo "c4 share rc=$?"
timeout 900 python tools/run_config.py --config c3 c5a c5b --out gpurun_out/final_configs.jsonl > /dev/null 2>> gpurun_out/final_configs.err; echo "c3 c5 rc=$?"
timeout 900 python tools/run_config.py --config c3 --fused 0 --out gpurun_out/final_configs.jsonl > /dev/null 2>> gpurun_out/final_configs.err; echo "c3 unfused rc=$?"
timeout 600 python bench.py --cpu-ops > gpurun_out/cpu_ops_r02.json 2> gpurun_out/cpu_ops.err; echo "cpu-ops rc=$?"
