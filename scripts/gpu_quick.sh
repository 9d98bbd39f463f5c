# quick GPU check: all parity tests, the bench's DQN leg, warm per-kernel durations of the TD update
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-aux > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_quick.json'))
print("env", d["value"], "e2e", d["e2e"]["value"]); print("dqn us/update", d["dqn"]["us_per_update"], "frac", d["dqn"]["roofline"]["frac"], "selfplay", d["dqn"]["selfplay_eps_greedy_steps_per_s"])
PY
tail -3 gpurun_out/bench_quick.err
timeout 600 ncu --metrics gpu__time_duration.sum,sm__cycles_active.avg --cache-control none --clock-control none -k regex:"l0_pair|l1_gemm|td_delta|dw_gemm" -s 12 -c 16 --csv --log-file gpurun_out/launches_td_warm.csv python scripts/td_only.py > gpurun_out/ncu_td_warm.log 2>&1
python - <<'PY'
import csv, collections
rows=list(csv.reader(open('gpurun_out/launches_td_warm.csv')))
hdr=[r for r in rows if 'Kernel Name' in r][0]
ik,im,iv=hdr.index('Kernel Name'),hdr.index('Metric Name'),hdr.index('Metric Value')
agg=collections.defaultdict(list)
for r in rows:
    if len(r)==len(hdr) and r[0].isdigit(): agg[(r[ik].split('(')[0][:40],r[im])].append(float(r[iv].replace(',','')))
for k,v in sorted(agg.items()): print(f"{k[0]:42s} {k[1]:28s} n={len(v):3d} mean={sum(v)/len(v):10.1f} min={min(v):10.1f}")
PY
