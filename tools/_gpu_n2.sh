set -x
cd $GRAFT_REPO_ROOT
export ITSOLV_BACKTRACE=1
nvidia-smi --query-gpu=index,name,memory.total --format=csv
python -X faulthandler -m pytest tests -m gpu -q > gpurun_out/r2_pytest_n2.log 2>&1; tail -6 gpurun_out/r2_pytest_n2.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; tail -c 2500 gpurun_out/r2_bench_n2.json; tail -15 gpurun_out/r2_bench_n2.err
