# round 2, second evidence pass on one B200 (after the piece tables / memory views / API-mode kernels): tests, smoke, bench (+ reference arm),
# ncu launch list of the bench command, ncu --set full of the API-mode kernels
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_r2b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r2b.log; tail -4 gpurun_out/pytest_gpu_r2b.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_bench_1gpu.json 2> gpurun_out/r2b_bench_1gpu.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2b_bench_reference_arm.json 2> gpurun_out/r2b_bench_ref.err; echo "ref rc=$?"
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2b_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$?"
timeout 300 python scripts/ncu_lane.py > gpurun_out/ncu_lane_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"legal_moves_lane_kernel|legal_moves_team_kernel|step_kernel|pick_random" -s 2 -c 3 -f -o gpurun_out/r2b_api python scripts/ncu_lane.py > gpurun_out/ncu_api.log 2>&1; echo "ncu api rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r2b_bench_1gpu.json')); q=d['dqn']
print('value %.3e e2e %.3e serial %.3e' % (d['value'], d['e2e']['value'], d['e2e']['serial_value']), d['timing']); print('api', {k:(v['steps_per_s'],v['us_per_ply']) for k,v in d['aux']['api_mode'].items()}, 'c5 %.3e' % d['aux']['config5_1M_envs_steps_per_s'], 'traced %.3e' % d['aux']['traced_steps_per_s'])
print('td', q['us_per_update'], q['us_per_update_calls'], 'selfplay %.3e' % q['selfplay_eps_greedy_steps_per_s'], q['train_loop']['ms_per_round'])"
cat gpurun_out/r2b_bench_reference_arm.json | cut -c1-300
