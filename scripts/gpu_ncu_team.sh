# ncu --set full of the team rollout kernel at 4096 envs (XQ_TEAM_MINB=1) and at ${BIG:-262144} envs (XQ_TEAM_MINB=8)
cd $GRAFT_REPO_ROOT
XQ_ROLLOUT_TEAM=4 XQ_TEAM_MINB=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout_ -s 3 -c 1 -f -o gpurun_out/prof_team4_4096 python bench.py --envs 4096 --steps 2 --warmup 3 --no-cpu-baseline --no-dqn --no-aux > gpurun_out/ncu_team4_4096.log 2>&1; echo "ncu rc=$?"
XQ_ROLLOUT_TEAM=4 XQ_TEAM_MINB=8 timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout_ -s 3 -c 1 -f -o gpurun_out/prof_team4_big python bench.py --envs ${BIG:-262144} --steps 2 --warmup 3 --no-cpu-baseline --no-dqn --no-aux > gpurun_out/ncu_team4_big.log 2>&1; echo "ncu rc=$?"
