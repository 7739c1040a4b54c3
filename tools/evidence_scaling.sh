#!/bin/bash
# the driver's scaling run in small: the bench line at N = 8 (with the configurations at their stated sizes), 4 and 2,
# per-rank step records, and what the host link delivers to N GPUs at once
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
tr() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29$((RANDOM % 800 + 100)) "$@"; }
ITSOLV_BENCH_RANK_DETAIL=gpurun_out/scale_n8 timeout 900 bash -c "$(declare -f tr); tr 8 bench.py --gpus 8 --steps 20 --warmup 5" > gpurun_out/scale_n8.json 2> gpurun_out/scale_n8.err; echo "n8 rc=$?"
for n in 4 2; do
  CUDA_VISIBLE_DEVICES=$(seq -s, 0 $((n-1))) timeout 600 bash -c "$(declare -f tr); tr $n bench.py --gpus $n --steps 20 --warmup 5 --no-configs" > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err; echo "n$n rc=$?"
done
for n in 8 4 2 1; do
  CUDA_VISIBLE_DEVICES=$(seq -s, 0 $((n-1))) timeout 200 bash -c "$(declare -f tr); tr $n tools/h2d_rates.py" 2>/dev/null | grep '^{' >> gpurun_out/h2d_rates.jsonl
done
cat gpurun_out/h2d_rates.jsonl | cut -c1-120
