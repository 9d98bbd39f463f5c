cd $GRAFT_REPO_ROOT
XQ_NVCC_EXTRA="-DXQ_TIMELINE" python cn_chess_ai_b200/build.py -f > /dev/null 2>&1
timeout 300 python scripts/tl_rollout.py 2>&1 | tee gpurun_out/timeline_rollout.txt | tail -22
