cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_selfplay_gpu.py tests/test_trainer_gpu.py tests/test_adapter_gpu.py tests/test_dqn_gpu.py -m gpu -q -x 2>&1 | tail -5
for L in 1 0; do XQ_ACT_LANE=$L timeout 300 python scripts/selfplay_probe.py; done
XQ_ACT_LANE=1 XQ_COLLECT_STREAMS=1 timeout 300 python scripts/selfplay_probe.py
XQ_ACT_LANE=1 XQ_COLLECT_STREAMS=3 timeout 300 python scripts/selfplay_probe.py
