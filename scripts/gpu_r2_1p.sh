cd $GRAFT_REPO_ROOT
timeout 300 python scripts/td_only.py > gpurun_out/td_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"act_team|q90_gemm" -s 12 -c 4 -f -o gpurun_out/r2b_act python scripts/td_only.py > gpurun_out/ncu_act2.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_act2.log
