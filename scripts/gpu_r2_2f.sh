cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
for Z in 1 0; do XQ_IO_ZEROCOPY=$Z timeout 600 python bench.py --steps 20 --warmup 5 --no-dqn --no-aux --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('zerocopy=$Z value %.4e e2e %.4e' % (d['value'], d['e2e']['value']))"; done
XQ_TRAIN_PROFILE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --no-aux 2>&1 | grep "xq_train_run rank" | tail -2
