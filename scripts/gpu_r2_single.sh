cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -8 gpurun_out/pytest_gpu.log
for T in 1 4; do XQ_ROLLOUT_TEAM=$T timeout 300 python scripts/rollout_sweep.py; done
XQ_ROLLOUT_TEAM=1 timeout 300 python scripts/rollout_sweep.py --traced
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
