set -x
cd $GRAFT_REPO_ROOT
export ITSOLV_BACKTRACE=1
python -X faulthandler -m pytest tests -m gpu -q > gpurun_out/r2_pytest2.log 2>&1; tail -15 gpurun_out/r2_pytest2.log
rm -f gpurun_out/r2_configs2.jsonl gpurun_out/r2_memtable2.jsonl
timeout 600 python tools/run_config.py --memory-table --out gpurun_out/r2_memtable2.jsonl 2>&1 | tail -15
timeout 600 python tools/run_config.py --config c3 --fused 0 --out gpurun_out/r2_configs2.jsonl 2>&1 | tail -3
timeout 600 python tools/run_config.py --config c5a --max-p 4 --n 2e7 --out gpurun_out/r2_configs2.jsonl 2>&1 | tail -3
timeout 600 python tools/run_config.py --config c5b --out gpurun_out/r2_configs2.jsonl 2>&1 | tail -3
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err; tail -c 6000 gpurun_out/r2_bench2.json; tail -5 gpurun_out/r2_bench2.err
tools/fp64_peak > gpurun_out/r2_fp64_peak.json; cat gpurun_out/r2_fp64_peak.json
timeout 600 python tools/opbench.py --json gpurun_out/r2_opbench.json > gpurun_out/r2_opbench.txt 2>&1; tail -45 gpurun_out/r2_opbench.txt
