set -x
cd $GRAFT_REPO_ROOT
export ITSOLV_BACKTRACE=1
nvidia-smi --query-gpu=index,name,memory.total --format=csv | head -3; free -g | head -2; nproc
export ITSOLV_WORKER_PROGRESS=gpurun_out/worker8_r02
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 tests/multi_gpu_worker.py > gpurun_out/worker8_r02.log 2>&1; echo "worker8 peer rc=$?"; grep "ok {" gpurun_out/worker8_r02.log | cut -c1-200 | head -8; tail -3 gpurun_out/worker8_r02.log | cut -c1-300
ITSOLV_P2P_ALLREDUCE=-1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 tests/multi_gpu_worker.py > gpurun_out/worker8_nccl_r02.log 2>&1; echo "worker8 nccl rc=$?"; grep -c "ok {" gpurun_out/worker8_nccl_r02.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; echo "bench8 rc=$?"; tail -c 1500 gpurun_out/r2_bench_n8.json; tail -5 gpurun_out/r2_bench_n8.err | cut -c1-300
CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29524 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r2_bench_n4.json 2> gpurun_out/r2_bench_n4.err; echo "bench4 rc=$?"; tail -c 800 gpurun_out/r2_bench_n4.json
