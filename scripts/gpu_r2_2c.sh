cd $GRAFT_REPO_ROOT
run() { n=$1; shift; env "$@" XQ_PROBE_CALLS=6 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29551 scripts/dist_td_probe.py 2>&1 | grep "^world"; }
for n in 1 2; do
run $n XQ_X=1
run $n XQ_PROBE_MAIN_PRIO=1
run $n XQ_PROBE_MAIN_PRIO=1 XQ_TD_GEMM_SPLITS=4
run $n XQ_PROBE_MAIN_PRIO=1 XQ_TD_GEMM_SPLITS=6
done
