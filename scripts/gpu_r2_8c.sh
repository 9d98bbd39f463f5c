cd $GRAFT_REPO_ROOT
N=${NGPU:-8}
run() { env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 scripts/dist_td_probe.py 2>&1 | grep "^world"; }
run XQ_X=1
run XQ_TD_GEMM_SPLITS=4
run XQ_PROBE_MAIN_PRIO=1
run XQ_TD_EARLY_GEMM=1
run XQ_DIST_FUSED_MODE=allgather
nproc
