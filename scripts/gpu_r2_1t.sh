cd $GRAFT_REPO_ROOT
run() { env "$@" XQ_PROBE_CALLS=10 timeout 300 python scripts/dist_td_probe.py 2>&1 | grep "^world\|rror" | cut -c1-200; }
run XQ_TD_FIRST_LATE=0
run XQ_TD_FIRST_LATE=1
run XQ_TD_FIRST_LATE=2
run XQ_TD_FIRST_LATE=4
run XQ_TD_FIRST_LATE=0
