# acting path A/B: XQ_ACT_CARRY=0 (layer-0 kernel every ply) vs 1 (sums carried by the act kernel's tail); parity tests first; per-kernel launch list
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests/test_selfplay_gpu.py tests/test_trainer_gpu.py tests/test_adapter_gpu.py -q -x > gpurun_out/pytest_selfplay.log 2>&1; tail -4 gpurun_out/pytest_selfplay.log
for T in 1 0; do
  XQ_ACT_CARRY=$T timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-aux > gpurun_out/bench_inc$T.json 2> gpurun_out/bench_inc$T.err; echo "bench rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/bench_inc$T.json')); print('carry $T: selfplay', d['dqn']['selfplay_eps_greedy_steps_per_s'], 'td us', d['dqn']['us_per_update'])"
done
for T in 1 0; do
XQ_ACT_CARRY=$T timeout 600 ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none -k regex:"q90|act_" -s 8 -c 30 --csv --log-file gpurun_out/launches_selfplay$T.csv python scripts/td_only.py > gpurun_out/ncu_selfplay.log 2>&1
python - <<PY
import csv, collections
rows=list(csv.reader(open('gpurun_out/launches_selfplay$T.csv')))
hdr=[r for r in rows if 'Kernel Name' in r][0]
ik,im,iv=hdr.index('Kernel Name'),hdr.index('Metric Name'),hdr.index('Metric Value')
agg=collections.defaultdict(list)
for r in rows:
    if len(r)==len(hdr) and r[0].isdigit(): agg[(r[ik].split('(')[0][:40],r[im])].append(float(r[iv].replace(',','')))
for k,v in sorted(agg.items()): print(f"{k[0]:42s} {k[1]:28s} n={len(v):3d} mean={sum(v)/len(v):10.1f} min={min(v):10.1f}")
PY
done
