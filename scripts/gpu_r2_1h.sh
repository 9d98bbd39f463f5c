cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { env "$@" XQ_PROBE_CALLS=8 timeout 300 python scripts/dist_td_probe.py 2>&1 | grep "^world\|Error\|error" ; }
run XQ_TD_EARLY_GEMM=5
run XQ_TD_EARLY_GEMM=2
run XQ_TD_EARLY_GEMM=5 XQ_TD_GEMM_SPLITS=4
run XQ_TD_EARLY_GEMM=5 XQ_TD_GEMM_SPLITS=6
run XQ_TD_EARLY_GEMM=5 XQ_TD_MAIN_PRIO=1
timeout 600 python -m pytest tests/test_env_gpu.py -x -q -m gpu -k "io" 2>&1 | tail -3
XQ_TD_EARLY_GEMM=5 timeout 600 python -m pytest tests/test_dqn_fast_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py --no-dqn --no-aux --no-cpu-baseline > gpurun_out/r2h_bench_e2e.json 2> gpurun_out/r2h_bench_e2e.err; tail -c 600 gpurun_out/r2h_bench_e2e.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2h_bench_e2e.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'serial',d['e2e']['serial_value'])
PY
