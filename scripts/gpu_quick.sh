# quick GPU check: DQN + selfplay parity tests and the bench's DQN leg
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-aux > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_quick.json'))
print("env", d["value"], "e2e", d["e2e"]["value"]); print("dqn", {k:v for k,v in d["dqn"].items() if k!="roofline"}); print(d["dqn"]["roofline"]["frac"])
PY
tail -3 gpurun_out/bench_quick.err
