# A/B of the fused rollout kernels: threads per board 16 (rollout_slots_kernel) vs 4 / 8 (rollout_team_kernel); parity tests, then throughput
cd $GRAFT_REPO_ROOT
for T in ${TEAMS:-4 8 16}; do
  echo "== XQ_ROLLOUT_TEAM=$T"
  XQ_ROLLOUT_TEAM=$T timeout 900 python -m pytest tests/test_env_gpu.py -q -x 2>&1 | tail -2
  XQ_ROLLOUT_TEAM=$T timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-dqn > gpurun_out/bench_team$T.json 2> gpurun_out/bench_team$T.err; echo "bench rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/bench_team$T.json')); print('team $T: 4096 envs', d['value'], 'e2e', d['e2e']['value'], '1M envs', d['aux']['config5_1M_envs_steps_per_s'])"
done
