cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_env_gpu.py tests/test_selfplay_gpu.py -x -q -m gpu 2>&1 | tail -3
XQ_ROLLOUT_TEAM=4 timeout 300 python scripts/rollout_sweep.py 2>&1 | tail -6
XQ_ROLLOUT_TEAM=1 timeout 300 python scripts/rollout_sweep.py 2>&1 | tail -6
timeout 300 python scripts/api_sweep.py 2>&1 | grep envs=
