# round 2 evidence pass on one B200: tests, smoke, bench (+ reference arm), ncu launch list of the bench command, ncu --set full of the API-mode and TD kernels
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_ref.err; echo "ref rc=$?"
# launch list of the bench command (per-launch times are cold-cache and serialised: shares, not absolutes)
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$?"
# full captures (each program ran clean above / in earlier calls of this round)
timeout 300 python scripts/ncu_lane.py > gpurun_out/ncu_lane_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"legal_moves_lane_kernel|step_kernel|pick_random" -s 2 -c 3 -f -o gpurun_out/r2_api python scripts/ncu_lane.py > gpurun_out/ncu_api.log 2>&1; echo "ncu api rc=$?"
timeout 300 python scripts/td_only.py > gpurun_out/td_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"l0_pair|l1_gemm|td_delta|dw_gemm" -s 8 -c 8 -f -o gpurun_out/r2_td python scripts/td_only.py > gpurun_out/ncu_td.log 2>&1; echo "ncu td rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_1gpu.json')); q=d['dqn']
print('value %.3e e2e %.3e' % (d['value'], d['e2e']['value']), d['timing']); print('td', q['us_per_update'], q['us_per_update_calls'], 'selfplay %.3e' % q['selfplay_eps_greedy_steps_per_s'], q['train_loop']['ms_per_round'])"
cat gpurun_out/r2_bench_reference_arm.json | cut -c1-400
