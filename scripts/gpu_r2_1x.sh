cd $GRAFT_REPO_ROOT
timeout 300 python scripts/ncu_lane.py > gpurun_out/ncu_lane_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"rollout_lane_kernel|legal_moves_lane_kernel" -s 3 -c 2 -f -o gpurun_out/r2c_lane python scripts/ncu_lane.py > gpurun_out/ncu_lane3.log 2>&1; echo "ncu rc=$?"; tail -1 gpurun_out/ncu_lane3.log
