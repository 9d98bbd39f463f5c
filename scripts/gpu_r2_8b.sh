cd $GRAFT_REPO_ROOT
N=${NGPU:-8}
timeout 900 python -m pytest tests/test_dist_gpu.py -q -x > gpurun_out/pytest_dist_${N}b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_dist_${N}b.log; tail -6 gpurun_out/pytest_dist_${N}b.log
XQ_LIB_PATH=$GRAFT_REPO_ROOT/cn_chess_ai_b200/libxq_b200_tl.so timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 scripts/dist_timeline.py > gpurun_out/dist_timeline_${N}.txt 2> gpurun_out/dist_timeline_${N}.err; echo "timeline rc=$?"; grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/dist_timeline_${N}.txt | head -100
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --no-aux > gpurun_out/bench_${N}gpu_b.json 2> gpurun_out/bench_${N}gpu_b.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_${N}gpu_b.json')); q=d['dqn']
    print('td us', q['us_per_update'], q.get('us_per_update_median'), q.get('us_per_update_min'), 'selfplay %.3e' % q['selfplay_eps_greedy_steps_per_s'], json.dumps(q.get('train_loop'))[-300:])
except Exception as e: print('parse failed', e)
PY
tail -3 gpurun_out/bench_${N}gpu_b.err
