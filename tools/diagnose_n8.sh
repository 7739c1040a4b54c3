#!/bin/bash
# where does a step at N=8 spend its time: the same bench line under five settings, with per-rank per-step records
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() {
  tag=$1; shift
  env "$@" ITSOLV_BENCH_RANK_DETAIL=gpurun_out/diag_$tag timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 \
    --master-addr 127.0.0.1 --master-port 29$((RANDOM % 800 + 100)) bench.py --gpus 8 --steps 20 --warmup 5 --no-e2e \
    --no-cpu-baseline --no-parity --no-configs --no-other-path > gpurun_out/diag_$tag.json 2> gpurun_out/diag_$tag.err
  echo "diag $tag rc=$?"
}
nvidia-smi topo -m > gpurun_out/topo_n8.txt 2>&1
lscpu | head -30 > gpurun_out/lscpu_n8.txt 2>&1
run A X=1
run B ITSOLV_BENCH_BIND=0
run C ITSOLV_P2P_HALO=-1
run D ITSOLV_P2P_ALLREDUCE=-1
run E ITSOLV_P2P_TIMEOUT_S=0
run F X=1
