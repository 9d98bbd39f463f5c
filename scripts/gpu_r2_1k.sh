cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_env_gpu.py -x -q -m gpu 2>&1 | tail -3
XQ_ROLLOUT_TEAM=1 timeout 300 python scripts/rollout_sweep.py 2>&1 | tail -4
XQ_LIB_PATH=$GRAFT_REPO_ROOT/cn_chess_ai_b200/libxq_b200_lb6.so XQ_ROLLOUT_TEAM=1 timeout 300 python scripts/rollout_sweep.py 2>&1 | tail -4
XQ_ROLLOUT_TEAM=4 XQ_TEAM_VIEW=1 timeout 300 python scripts/rollout_sweep.py 2>&1 | tail -4 | head -2
