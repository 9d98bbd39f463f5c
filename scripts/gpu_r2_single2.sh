cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -8 gpurun_out/pytest_gpu.log
for T in 1 0; do XQ_LEGAL_LANE=$T timeout 300 python scripts/api_sweep.py; done
timeout 300 python scripts/ncu_lane.py > gpurun_out/ncu_lane_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"rollout_lane_kernel|rollout_team_kernel|legal_moves_lane_kernel|step_kernel" -s 4 -c 4 -o gpurun_out/r2_lane python scripts/ncu_lane.py > gpurun_out/ncu_lane.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_lane.log
