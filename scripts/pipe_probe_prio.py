import os, sys
sys.path.insert(0, "/root/repo")
import torch
import cn_chess_ai_b200 as xq
prio = int(os.environ.get("PRIO", "0"))
s = torch.cuda.Stream(priority=prio)
with torch.cuda.stream(s):
    env = xq.BatchedEnv(65536, seed=1); net = xq.DQN(lr=1e-6); rb = xq.ReplayBuffer(1 << 20)
    env.set_stream(s.cuda_stream); net.set_stream(s.cuda_stream)
    xq.collect(net, env, rb, 16, 0.1)
    N = 64
    xq.td_update_replay_n(net, rb, 4096, 5, 0, 8, True, 1e-6)
    best = 1e9
    for _ in range(4):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(s); xq.td_update_replay_n(net, rb, 4096, 5, 100, N, True, 1e-6); b.record(s); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    print("prio", prio, "pipelined us/update", 1e3 * best / N)
