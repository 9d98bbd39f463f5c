# multi-GPU check (gpurun --gpus N): the 2-GPU exchange test, then the bench as the driver launches it
cd $GRAFT_REPO_ROOT
N=${NGPU:-2}
timeout 600 python -m pytest tests/test_dist_gpu.py -q -x 2>&1 | tail -15
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "bench rc=$?"
cat gpurun_out/bench_${N}gpu.json | python -c "
import json,sys; d=json.load(sys.stdin); print('n_gpus', d['n_gpus'], 'value', d['value'], 'e2e', d['e2e']['value']); q=d['dqn']; print('td us', q['us_per_update'], q['grad_allreduce'], 'selfplay', q['selfplay_eps_greedy_steps_per_s'])"
tail -3 gpurun_out/bench_${N}gpu.err
