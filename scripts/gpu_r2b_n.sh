# round 2, second N-GPU evidence pass (gpurun --gpus N; after the piece tables / memory views / API-mode kernels): the dist test on every GPU, then the bench exactly as the driver launches it
cd $GRAFT_REPO_ROOT
N=${NGPU:-8}
timeout 900 python -m pytest tests/test_dist_gpu.py -q -x > gpurun_out/pytest_dist_r2b_${N}.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_dist_r2b_${N}.log; tail -4 gpurun_out/pytest_dist_r2b_${N}.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2b_bench_${N}gpu.json 2> gpurun_out/r2b_bench_${N}gpu.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/r2b_bench_${N}gpu.json')); q=d['dqn']
print('n_gpus', d['n_gpus'], 'value %.3e' % d['value'], 'e2e %.3e' % d['e2e']['value'], 'traced %.3e' % d['aux']['traced_steps_per_s'], 'config5 %.3e' % d['aux']['config5_1M_envs_steps_per_s'], 'api', {k: '%.3e' % v['steps_per_s'] for k, v in d['aux']['api_mode'].items()})
print('   td us', q['us_per_update'], q['us_per_update_calls'], 'selfplay %.3e' % q['selfplay_eps_greedy_steps_per_s'], 'replicas', q['replicas_identical'])
print('   train_loop', json.dumps(q.get('train_loop'))[-430:])
PY
tail -2 gpurun_out/r2b_bench_${N}gpu.err
