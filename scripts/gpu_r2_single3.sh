cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -12 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python scripts/api_sweep.py
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); q=d['dqn']
print('value %.3e e2e %.3e' % (d['value'], d['e2e']['value'])); print('aux', json.dumps(d['aux'])[:1500]); print('td', q['us_per_update'], q['us_per_update_calls'], 'selfplay %.3e' % q['selfplay_eps_greedy_steps_per_s']); print(json.dumps(q.get('adapter_train')), json.dumps(q.get('reference_train'))); print(json.dumps(d['cpu_baseline'])[:1500])"; tail -3 gpurun_out/bench.err
