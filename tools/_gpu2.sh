set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2_pytest2.log; tail -8 gpurun_out/r2_pytest2.log
rm -f gpurun_out/r2_configs2.jsonl gpurun_out/r2_memtable2.jsonl
timeout 600 python tools/run_config.py --memory-table --out gpurun_out/r2_memtable2.jsonl 2>&1 | tail -15
timeout 600 python tools/run_config.py --config c3 --fused 0 --out gpurun_out/r2_configs2.jsonl 2>&1 | tail -3
timeout 600 python tools/run_config.py --config c5a --max-p 4 --n 2e7 --out gpurun_out/r2_configs2.jsonl 2>&1 | tail -3
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err; tail -c 3000 gpurun_out/r2_bench2.json; tail -5 gpurun_out/r2_bench2.err
