# throughput of the fused rollout kernels over the env count: 16 threads per board vs teams of 4 (register caps via XQ_TEAM_MINB)
# SIZES="4096 1048576" VARS="16:1 4:1 4:8" PARITY=1 bash scripts/gpu_team_sweep.sh
cd $GRAFT_REPO_ROOT
if [ -n "$PARITY" ]; then for T in 4 16; do XQ_ROLLOUT_TEAM=$T timeout 900 python -m pytest tests/test_env_gpu.py -q -x 2>&1 | tail -1; done; fi
for N in ${SIZES:-4096 16384 65536 1048576}; do
  for V in ${VARS:-16:1 4:1 4:6 4:8}; do
    T=${V%%:*}; M=${V##*:}
    XQ_ROLLOUT_TEAM=$T XQ_TEAM_MINB=$M timeout 300 python bench.py --envs $N --steps 8 --warmup 3 --no-cpu-baseline --no-dqn --no-aux > gpurun_out/sweep.json 2> gpurun_out/sweep.err || tail -3 gpurun_out/sweep.err
    python -c "
import json; d=json.load(open('gpurun_out/sweep.json')); print('envs $N team $T minb $M: %.3e steps/s  e2e %.3e' % (d['value'], d['e2e']['value']))"
  done
done
