cd $GRAFT_REPO_ROOT
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
cat > /tmp/td_only.py <<'PY'
import torch, cn_chess_ai_b200 as xq
s = torch.cuda.current_stream()
env = xq.BatchedEnv(65536, seed=1); net = xq.DQN(lr=1e-6); rb = xq.ReplayBuffer(1 << 20)
env.set_stream(s.cuda_stream); net.set_stream(s.cuda_stream)
xq.collect(net, env, rb, 16, 0.1)
for i in range(8): xq.td_update_replay(net, rb, 4096, 5, i, True, 1e-6)
torch.cuda.synchronize()
PY
timeout 300 python /tmp/td_only.py > gpurun_out/plain_td.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 80 --csv --log-file gpurun_out/launches_td.csv python /tmp/td_only.py > gpurun_out/ncu_td.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/launches_td.csv')) if len(r)>10 and r[0].isdigit()]
agg=collections.defaultdict(list)
for r in rows: agg[r[4].split('(')[0][:60]].append(float(r[-1]))
for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1])): print(f"{k:62s} n={len(v):3d} mean={sum(v)/len(v)/1e3:9.2f} us")
PY
