cd $GRAFT_REPO_ROOT
timeout -s KILL 600 python -m pytest tests/test_dqn_fast_gpu.py tests/test_selfplay_gpu.py tests/test_adapter_gpu.py -x -q > gpurun_out/pytest_fast.log 2>&1; echo rc=$?; tail -15 gpurun_out/pytest_fast.log
timeout 300 python scripts/td_only.py > gpurun_out/plain_td.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 80 --csv --log-file gpurun_out/launches_td.csv python scripts/td_only.py > gpurun_out/ncu_td.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/launches_td.csv')) if len(r)>10 and r[0].isdigit()]
agg=collections.defaultdict(list)
for r in rows: agg[r[4].split('(')[0][:60]].append(float(r[-1]))
tot=0
for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1])): print(f"{k:62s} n={len(v):3d} mean={sum(v)/len(v)/1e3:9.2f} us")
PY
