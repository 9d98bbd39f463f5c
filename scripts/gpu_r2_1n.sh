cd $GRAFT_REPO_ROOT
timeout 300 python scripts/ncu_lane.py > gpurun_out/ncu_lane_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"rollout_lane_kernel|rollout_team_kernel" -s 4 -c 2 -f -o gpurun_out/r2b_lane python scripts/ncu_lane.py > gpurun_out/ncu_lane2.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_lane2.log
