"""the multi-GPU update path on ONE GPU (world = 1, connected to itself): device and CPU-enqueue time per update of the pipelined call,
to separate the cost of the extra exchange kernel from the cross-GPU wait"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cn_chess_ai_b200 as xq
s = torch.cuda.current_stream()
env = xq.BatchedEnv(65536, seed=1); net = xq.DQN(lr=1e-6); rb = xq.ReplayBuffer(1 << 20)
env.set_stream(s.cuda_stream); net.set_stream(s.cuda_stream)
xq.collect(net, env, rb, 16, 0.1)
if os.environ.get("XQ_PROBE_CONNECT", "1") != "0":
    net.dist_connect(0, 1, net.dist_export().reshape(1, 64))
N = 60
xq.td_update_replay_n(net, rb, 4096, 5, 0, 8, True, 1e-6)
best = 1e9
for _ in range(3):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(s); xq.td_update_replay_n(net, rb, 4096, 5, 100, N, True, 1e-6); b.record(s); torch.cuda.synchronize()
    best = min(best, a.elapsed_time(b))
print("pipelined us/update", 1e3 * best / N)
torch.cuda.synchronize()
t0 = time.perf_counter(); xq.td_update_replay_n(net, rb, 4096, 5, 300, N, True, 1e-6); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print("CPU enqueue us/update", 1e6 * (t1 - t0) / N, " total us/update", 1e6 * (t2 - t0) / N, "timed out", net.dist_timed_out())
