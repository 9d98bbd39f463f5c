cd $GRAFT_REPO_ROOT
export XQ_SWEEP_SIZES=4096:100,8192:100,12288:100,16384:100,24576:100,32768:100
XQ_LEGAL_TEAM=1 timeout 300 python scripts/api_sweep.py 2>&1 | grep envs= | sed 's/; step.*//'
XQ_LEGAL_TEAM=0 timeout 300 python scripts/api_sweep.py 2>&1 | grep envs= | sed 's/; step.*//'
