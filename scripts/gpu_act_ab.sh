# A/B of the self-play act kernels: team of 4 threads per board (act_team_kernel, default) vs the generic thread-per-board kernel (XQ_ACT_TEAM=0)
cd $GRAFT_REPO_ROOT
for T in 1 0; do
  echo "== XQ_ACT_TEAM=$T"
  XQ_ACT_TEAM=$T timeout 900 python -m pytest tests/test_selfplay_gpu.py tests/test_trainer_gpu.py tests/test_adapter_gpu.py -q -x 2>&1 | tail -2
  XQ_ACT_TEAM=$T timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-aux > gpurun_out/bench_act$T.json 2> gpurun_out/bench_act$T.err; echo "bench rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/bench_act$T.json')); print('act team $T: selfplay', d['dqn']['selfplay_eps_greedy_steps_per_s'], 'td us', d['dqn']['us_per_update'])"
done
