# ncu --set full of one launch of the self-play act kernel (65,536 envs, scripts/td_only.py)
cd $GRAFT_REPO_ROOT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"act_team" -s 6 -c 1 -f -o gpurun_out/prof_act_team python scripts/td_only.py > gpurun_out/ncu_act_team.log 2>&1; echo "ncu rc=$?"
