cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none -k regex:"q90|act_" -s 8 -c 30 --csv --log-file gpurun_out/launches_selfplay.csv python scripts/td_only.py > gpurun_out/ncu_selfplay.log 2>&1
python - <<PY
import csv, collections
rows=list(csv.reader(open('gpurun_out/launches_selfplay.csv')))
hdr=[r for r in rows if 'Kernel Name' in r][0]
ik,im,iv=hdr.index('Kernel Name'),hdr.index('Metric Name'),hdr.index('Metric Value')
agg=collections.defaultdict(list)
for r in rows:
    if len(r)==len(hdr) and r[0].isdigit(): agg[(r[ik].split('(')[0][:40],r[im])].append(float(r[iv].replace(',','')))
for k,v in sorted(agg.items()): print(f"{k[0]:42s} {k[1]:28s} n={len(v):3d} mean={sum(v)/len(v):10.1f} min={min(v):10.1f}")
PY
