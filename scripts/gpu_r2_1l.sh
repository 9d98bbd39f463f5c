cd $GRAFT_REPO_ROOT
export XQ_SWEEP_SIZES=4096:200,6144:200,8192:200,10240:200,12288:200
XQ_ROLLOUT_TEAM=1 timeout 300 python scripts/rollout_sweep.py 2>&1 | tail -5
XQ_ROLLOUT_TEAM=4 XQ_TEAM_VIEW=0 timeout 300 python scripts/rollout_sweep.py 2>&1 | tail -5
XQ_ROLLOUT_TEAM=4 XQ_TEAM_VIEW=1 timeout 300 python scripts/rollout_sweep.py 2>&1 | tail -5
