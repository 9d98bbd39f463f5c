cd $GRAFT_REPO_ROOT
for L in 1 0; do XQ_ACT_LANE=$L timeout 300 python scripts/selfplay_probe.py; done
timeout 300 python scripts/td_only.py > gpurun_out/act_plain.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:"act_lane" -s 6 -c 2 -f -o gpurun_out/r2_act_lane python scripts/td_only.py > gpurun_out/ncu_act.log 2>&1; echo ncu rc=$?
