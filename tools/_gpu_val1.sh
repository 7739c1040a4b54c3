#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/val1_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/val1_pytest.log
timeout 900 python bench.py > gpurun_out/val1_bench.json 2> gpurun_out/val1_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/val1_bench.json
