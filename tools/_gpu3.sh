set -x
cd $GRAFT_REPO_ROOT
export ITSOLV_BACKTRACE=1
python -X faulthandler -m pytest tests -m gpu -q > gpurun_out/r2_pytest3.log 2>&1; tail -6 gpurun_out/r2_pytest3.log
rm -f gpurun_out/r2_configs3.jsonl
timeout 900 python tools/run_config.py --config c4 --n 2.5e8 --out gpurun_out/r2_configs3.jsonl 2>&1 | tail -3
timeout 600 python tools/run_config.py --config c3 --out gpurun_out/r2_configs3.jsonl 2>&1 | tail -3
# ncu: (1) the tile residual kernel, (2) wide Gram panels with the DMMA and with the FMA consumer, (3) launch list of one bench step
A='python tools/opbench.py --reps 2 --only davidson_residual[24x16],davidson_residual[40x16]'
$A > gpurun_out/plain_a.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:davidson_residual_tile -s 2 -c 2 -o gpurun_out/ncu_tile_r02 $A > gpurun_out/ncu_a.log 2>&1
B='python tools/opbench.py --reps 2 --only gemm_inner[16x64],gemm_inner[64x64]'
$B > gpurun_out/plain_b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_inner -s 4 -c 2 -o gpurun_out/ncu_gi_mma_r02 $B > gpurun_out/ncu_b.log 2>&1
C='python tools/opbench.py --reps 2 --opt GI_MMA=-1 --only gemm_inner[16x64],gemm_inner[64x64]'
$C > gpurun_out/plain_c.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_inner -s 4 -c 2 -o gpurun_out/ncu_gi_fma_r02 $C > gpurun_out/ncu_c.log 2>&1
D='python bench.py --steps 1 --warmup 1 --min-warmup 1 --no-e2e --no-cpu-baseline --no-parity --no-configs --no-other-path'
$D > gpurun_out/plain_d.log 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/ncu_dram_bench_r02.csv $D > gpurun_out/ncu_d.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/ncu_dram_bench_r02.csv; tail -3 gpurun_out/ncu_a.log gpurun_out/ncu_b.log gpurun_out/ncu_c.log gpurun_out/ncu_d.log
