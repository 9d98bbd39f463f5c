# ncu --set full of one warm launch of l0_act_kernel (65,536 envs, scripts/td_only.py): carried ply, and gather ply (XQ_ACT_INCREMENTAL=0)
cd $GRAFT_REPO_ROOT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"l0_act" -s 5 -c 1 -f -o gpurun_out/prof_l0act_inc python scripts/td_only.py > gpurun_out/ncu_acting.log 2>&1; echo "ncu rc=$?"
XQ_ACT_INCREMENTAL=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"l0_act" -s 5 -c 1 -f -o gpurun_out/prof_l0act_gather python scripts/td_only.py > gpurun_out/ncu_acting.log 2>&1; echo "ncu rc=$?"
