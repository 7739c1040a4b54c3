set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,memory.total,memory.used --format=csv
free -g | head -2; nproc
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2_pytest1.log; tail -5 gpurun_out/r2_pytest1.log
rm -f gpurun_out/r2_configs1.jsonl
timeout 600 python tools/run_config.py --config c3 --fused 0 --out gpurun_out/r2_configs1.jsonl 2>&1 | tail -3
timeout 600 python tools/run_config.py --config c3 --fused 1 --out gpurun_out/r2_configs1.jsonl 2>&1 | tail -3
timeout 600 python tools/run_config.py --config c5b --fused 0 --out gpurun_out/r2_configs1.jsonl 2>&1 | tail -3
timeout 600 python tools/run_config.py --config c5b --fused 1 --out gpurun_out/r2_configs1.jsonl 2>&1 | tail -3
timeout 600 python tools/run_config.py --config c5a --out gpurun_out/r2_configs1.jsonl 2>&1 | tail -3
timeout 600 python tools/run_config.py --config c5a --max-p 2 --out gpurun_out/r2_configs1.jsonl 2>&1 | tail -3
timeout 600 python tools/run_config.py --memory-table --out gpurun_out/r2_memtable.jsonl 2>&1 | tail -15
timeout 900 python bench.py > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; tail -c 1500 gpurun_out/r2_bench1.json
