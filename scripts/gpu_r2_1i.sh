cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_env_gpu.py -x -q -m gpu 2>&1 | tail -4
XQ_LEGAL_TEAM=1 timeout 300 python scripts/api_sweep.py 2>&1 | grep envs=
XQ_LEGAL_TEAM=0 timeout 300 python scripts/api_sweep.py 2>&1 | grep envs=
timeout 600 python bench.py --no-dqn --no-aux --no-cpu-baseline > gpurun_out/r2i_bench_e2e.json 2> gpurun_out/r2i_bench_e2e.err; tail -c 600 gpurun_out/r2i_bench_e2e.err
XQ_IO_SIDE_STREAMS=0 timeout 600 python bench.py --no-dqn --no-aux --no-cpu-baseline > gpurun_out/r2i_bench_e2e_noside.json 2> gpurun_out/r2i_bench_e2e_noside.err
python - <<'PY'
import json
for f in ('r2i_bench_e2e','r2i_bench_e2e_noside'):
    d=json.load(open(f'gpurun_out/{f}.json'))
    print(f,'value %.4e e2e %.4e serial %.4e'%(d['value'],d['e2e']['value'],d['e2e']['serial_value']))
PY
