cd $GRAFT_REPO_ROOT
run() { n=$1; shift; env "$@" XQ_PROBE_CALLS=6 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29551 scripts/dist_td_probe.py 2>&1 | grep "^world"; }
XQ_TD_EARLY_GEMM=4 timeout 600 python -m pytest tests/test_dqn_fast_gpu.py tests/test_dist_gpu.py tests/test_selfplay_gpu.py tests/test_trainer_gpu.py -m gpu -q -x 2>&1 | tail -3
for n in ${NS:-1 2}; do
run $n XQ_TD_EARLY_GEMM=2
run $n XQ_TD_EARLY_GEMM=4
run $n XQ_TD_EARLY_GEMM=4 XQ_TD_GEMM_A_TILES=16
run $n XQ_TD_EARLY_GEMM=4 XQ_TD_GEMM_A_TILES=24
run $n XQ_TD_EARLY_GEMM=4 XQ_TD_GEMM_A_TILES=28
done
