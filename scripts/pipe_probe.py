"""timing probe of the pipelined multi-update path (XQ_PIPE_DEBUG=0/1/2) against the sequential path"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cn_chess_ai_b200 as xq
s = torch.cuda.current_stream()
env = xq.BatchedEnv(65536, seed=1); net = xq.DQN(lr=1e-6); rb = xq.ReplayBuffer(1 << 20)
env.set_stream(s.cuda_stream); net.set_stream(s.cuda_stream)
xq.collect(net, env, rb, 16, 0.1)
def t(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(s); fn(); b.record(s); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best
N = 60
xq.td_update_replay_n(net, rb, 4096, 5, 0, 8, True, 1e-6)
print("pipelined us/update", 1e3 * t(lambda: xq.td_update_replay_n(net, rb, 4096, 5, 100, N, True, 1e-6)) / N)
def seq():
    for i in range(N): xq.td_update_replay(net, rb, 4096, 5, 100 + i, True, 1e-6)
print("sequential us/update", 1e3 * t(seq) / N)
import time
torch.cuda.synchronize()
t0 = time.perf_counter(); xq.td_update_replay_n(net, rb, 4096, 5, 300, N, True, 1e-6); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print("pipelined: CPU enqueue us/update", 1e6 * (t1 - t0) / N, " total us/update", 1e6 * (t2 - t0) / N)
t0 = time.perf_counter(); seq(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print("sequential: CPU enqueue us/update", 1e6 * (t1 - t0) / N, " total us/update", 1e6 * (t2 - t0) / N)
