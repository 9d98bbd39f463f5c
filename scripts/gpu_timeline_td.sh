cd $GRAFT_REPO_ROOT
XQ_NVCC_EXTRA="-DXQ_TIMELINE" python cn_chess_ai_b200/build.py -f > /dev/null 2>&1
timeout 600 python -m pytest tests/test_dqn_fast_gpu.py tests/test_selfplay_gpu.py -m gpu -q -x > gpurun_out/pytest_fast.log 2>&1; tail -3 gpurun_out/pytest_fast.log
timeout 300 python scripts/tl_dump.py > gpurun_out/timeline_s2.txt 2>&1; grep -v -E "^ +(1[2-9]|[4-9]|2[0-9]|3[0-5]) (MMA|EPI)" gpurun_out/timeline_s2.txt | tail -34
