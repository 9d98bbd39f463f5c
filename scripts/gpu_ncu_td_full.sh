cd $GRAFT_REPO_ROOT
timeout 300 python scripts/td_only.py > gpurun_out/plain_td.log 2>&1 && \
timeout 900 ncu --set full --cache-control none --clock-control none --import-source on -k regex:"l0_pair|l1_gemm|td_delta|dw_gemm" -s 16 -c 4 -f -o gpurun_out/prof_td_s2b python scripts/td_only.py > gpurun_out/ncu_td_full.log 2>&1
tail -3 gpurun_out/ncu_td_full.log
