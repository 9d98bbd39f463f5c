"""steps/s of the fused random-policy rollout for several env counts (device-resident, CUDA events); kernel chosen by XQ_ROLLOUT_TEAM"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cn_chess_ai_b200 as xq
from cn_chess_ai_b200._lib import check
s = torch.cuda.current_stream()
L = xq.lib()
traced = "--traced" in sys.argv
sizes = ((4096, 200), (16384, 200), (65536, 100), (1 << 20, 32))
if os.environ.get("XQ_SWEEP_SIZES"):      # e.g. "4096:200,8192:200"
    sizes = tuple(tuple(int(x) for x in t.split(":")) for t in os.environ["XQ_SWEEP_SIZES"].split(","))
for n, plies in sizes:
    env = xq.BatchedEnv(n, seed=7)
    env.set_stream(s.cuda_stream)
    run = (lambda: check(L.xq_env_rollout_random_traced_async(env.handle, plies, None))) if traced else (lambda: env.rollout_random_async(plies))
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    best = 1e9; tot = 0.0
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(s); run(); b.record(s)
        torch.cuda.synchronize()
        ms = a.elapsed_time(b); best = min(best, ms); tot += ms
    print(f"team={os.environ.get('XQ_ROLLOUT_TEAM','default')} traced={traced} envs={n} plies={plies} mean {n*plies*10/(tot*1e-3):.3e} best {n*plies/(best*1e-3):.3e} steps/s  ({tot/10:.3f} ms)")
    env.close()
