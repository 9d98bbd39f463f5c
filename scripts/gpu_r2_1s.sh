cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_env_gpu.py tests/test_selfplay_gpu.py tests/test_trainer_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python scripts/api_sweep.py 2>&1 | grep envs=
XQ_SWEEP_SIZES=4096:200,65536:100,1048576:32 timeout 300 python scripts/rollout_sweep.py 2>&1 | grep envs=
XQ_ACT_LANE=0 timeout 300 python scripts/selfplay_probe.py 2>&1 | grep envs= | head -1
