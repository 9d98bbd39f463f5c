cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_env_gpu.py -x -q -m gpu 2>&1 | tail -3
XQ_SWEEP_SIZES=32768:100,65536:50,1048576:10 timeout 300 python scripts/api_sweep.py 2>&1 | grep envs=
XQ_SWEEP_SIZES=16384:200,65536:100,1048576:32 timeout 300 python scripts/rollout_sweep.py 2>&1 | grep envs=
