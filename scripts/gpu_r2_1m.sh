cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_env_gpu.py -x -q -m gpu -k "board_per_thread or large_n or io or both_views" 2>&1 | tail -3
export XQ_SWEEP_SIZES=4096:200,16384:200,65536:100,1048576:32
XQ_ROLLOUT_TEAM=1 timeout 300 python scripts/rollout_sweep.py 2>&1 | tail -4
XQ_LIB_PATH=$GRAFT_REPO_ROOT/cn_chess_ai_b200/libxq_b200_lb5.so XQ_ROLLOUT_TEAM=1 timeout 300 python scripts/rollout_sweep.py 2>&1 | tail -4
