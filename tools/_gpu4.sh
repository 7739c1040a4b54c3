set -x
cd $GRAFT_REPO_ROOT
export ITSOLV_BACKTRACE=1
python -X faulthandler -m pytest tests -m gpu -q > gpurun_out/r2_pytest4.log 2>&1; tail -4 gpurun_out/r2_pytest4.log
rm -f gpurun_out/r2_configs4.jsonl
export ITSOLV_VERIFY_DETAIL=1
timeout 300 python tools/run_config.py --config c4 --n 2e6 --out gpurun_out/r2_configs4.jsonl 2>&1 | tail -2 | cut -c1-300
timeout 300 python tools/run_config.py --config c4 --n 2e7 --out gpurun_out/r2_configs4.jsonl 2>&1 | tail -2 | cut -c1-300
timeout 900 python tools/run_config.py --config c4 --n 2.5e8 --out gpurun_out/r2_configs4.jsonl 2>&1 | tail -2 | cut -c1-300
timeout 900 python tools/run_config.py --config c4 --n 2.5e8 --fused 0 --out gpurun_out/r2_configs4.jsonl 2>&1 | tail -2 | cut -c1-300
