# ncu evidence for the env rollout kernel: launch list + one full capture (run under gpurun, 1 GPU)
cd $GRAFT_REPO_ROOT
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-aux"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_rollout.csv $CMD > gpurun_out/ncu_list.log 2>&1
timeout 300 $CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:rollout_slots -s 3 -c 2 -o gpurun_out/prof_rollout_slots $CMD > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
