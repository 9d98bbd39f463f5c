# round 2, N-GPU check (gpurun --gpus N): the dist test on every GPU, then the bench as the driver launches it (owner mode), then only the TD leg in the all-gather mode
cd $GRAFT_REPO_ROOT
N=${NGPU:-8}
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_dist_gpu.py -q -x > gpurun_out/pytest_dist_${N}.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_dist_${N}.log; tail -15 gpurun_out/pytest_dist_${N}.log
XQ_TRAIN_PROFILE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${N}gpu_owner.json 2> gpurun_out/bench_${N}gpu_owner.err; echo "bench owner rc=$?"
XQ_DIST_FUSED_MODE=allgather timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline --no-aux > gpurun_out/bench_${N}gpu_allgather.json 2> gpurun_out/bench_${N}gpu_allgather.err; echo "bench allgather rc=$?"
python - <<PY
import json
for m in ("owner", "allgather"):
    try:
        d=json.load(open('gpurun_out/bench_${N}gpu_%s.json' % m)); q=d['dqn']
        print(m, 'n_gpus', d['n_gpus'], 'value %.3e' % d['value'], 'e2e %.3e' % d['e2e']['value'], 'aux', json.dumps(d.get('aux'))[:600])
        print('   td us', q['us_per_update'], q.get('us_per_update_median'), q.get('us_per_update_min'), 'selfplay %.3e' % q['selfplay_eps_greedy_steps_per_s'], json.dumps(q.get('train_loop'))[-420:])
    except Exception as e: print(m, 'parse failed', e)
PY
grep -h "xq_train_run rank" gpurun_out/bench_${N}gpu_owner.err | head -4
tail -3 gpurun_out/bench_${N}gpu_owner.err
