# evidence pass: tests, smoke, bench (+ reference arm), ncu launch list of the bench command, ncu --set full of the rollout and DQN kernels
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
if [ -z "$NO_REF" ]; then timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; fi
# launch list of the bench command (per-launch times are cold-cache and serialised: shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$?"
# full captures (each program ran clean above)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout_ -s 3 -c 1 -f -o gpurun_out/prof_rollout_final python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-dqn --no-aux > gpurun_out/ncu_rollout_final.log 2>&1; echo "ncu rollout rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout_ -s 3 -c 1 -f -o gpurun_out/prof_rollout_1m python bench.py --envs 1048576 --plies 32 --steps 2 --warmup 3 --no-cpu-baseline --no-dqn --no-aux > gpurun_out/ncu_rollout_1m.log 2>&1; echo "ncu rollout 1M rc=$?"
if [ -z "$NO_DQN" ]; then
  # acting path: a warm [q90_gemm -> act_team (with the carried layer-0 tail)] pair of the collector; TD update: the 4 kernels of two warm updates
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"q90_gemm|act_team" -s 20 -c 2 -f -o gpurun_out/prof_acting_final python scripts/td_only.py > gpurun_out/ncu_acting_final.log 2>&1; echo "ncu acting rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"l0_pair|l1_gemm|td_delta|dw_gemm" -s 8 -c 8 -f -o gpurun_out/prof_dqn_final python scripts/td_only.py > gpurun_out/ncu_dqn_final.log 2>&1; echo "ncu dqn rc=$?"
fi
