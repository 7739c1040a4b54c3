cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 200 python tools/opbench.py --only "select[4, diagonal]" --reps 2 > gpurun_out/prof_select_plain.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:select_pass -c 2 -f -o gpurun_out/prof_select \
  python tools/opbench.py --only "select[4, diagonal]" --reps 1 > gpurun_out/prof_select.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/prof_select.ncu-rep
