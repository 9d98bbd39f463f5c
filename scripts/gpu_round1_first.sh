set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
nproc
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench1.json 2> gpurun_out/bench1.err; echo "bench rc=$?"; cat gpurun_out/bench1.json; tail -5 gpurun_out/bench1.err
timeout 300 python tests/golden/make_golden.py --nn-gpu > gpurun_out/golden_nn.log 2>&1; tail -3 gpurun_out/golden_nn.log
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-aux > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-aux > gpurun_out/ncu_list.log 2>&1
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-aux > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:rollout_random -s 3 -c 2 -o gpurun_out/prof_rollout_r1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-aux > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
