# session 2, call A: full GPU parity tests, smoke, headline bench, ncu --set full of the five TD-update kernels
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_s2a.json 2> gpurun_out/bench_s2a.err; echo "bench rc=$?"; cat gpurun_out/bench_s2a.json; tail -5 gpurun_out/bench_s2a.err
timeout 300 python scripts/td_only.py > gpurun_out/plain_td.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"l0_pair|l1_gemm|td_delta|dw0_gemm|dw_reduce" -s 20 -c 5 -f -o gpurun_out/prof_td_s2a python scripts/td_only.py > gpurun_out/ncu_td_full.log 2>&1
tail -3 gpurun_out/ncu_td_full.log
