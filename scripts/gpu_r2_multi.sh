# round 2, multi-GPU check (gpurun --gpus N): all GPU tests (the dist test uses every GPU), then the bench at N and at 1
cd $GRAFT_REPO_ROOT
N=${NGPU:-2}
nvidia-smi -L | head -8
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_${N}.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_${N}.log; tail -25 gpurun_out/pytest_gpu_${N}.log
for MODE in owner allgather; do
XQ_DIST_FUSED_MODE=$MODE timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${N}gpu_$MODE.json 2> gpurun_out/bench_${N}gpu_$MODE.err; echo "bench $MODE rc=$?"
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_${N}gpu_$MODE.json')); print('n_gpus', d['n_gpus'], 'value %.3e' % d['value'], 'e2e %.3e' % d['e2e']['value'], d.get('timing'), d.get('aux')); q=d['dqn']
    print('td us', q['us_per_update'], q.get('us_per_update_median'), q.get('us_per_update_min'), q['grad_allreduce'], 'selfplay %.3e' % q['selfplay_eps_greedy_steps_per_s'], q.get('train_loop'))
except Exception as e: print('parse failed', e)
PY
tail -5 gpurun_out/bench_${N}gpu_$MODE.err
done
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo "bench1 rc=$?"; cat gpurun_out/bench_1gpu.json; tail -3 gpurun_out/bench_1gpu.err
