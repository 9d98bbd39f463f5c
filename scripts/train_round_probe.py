"""where the time of one xq_train_run round goes: wall clock of its parts (collect 4 plies / 64 TD updates / target sync / event drain), each synchronised"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import cn_chess_ai_b200 as xq  # noqa: E402
from cn_chess_ai_b200.trainer import drain_game_events, enable_game_events, train  # noqa: E402

s = torch.cuda.current_stream()
env = xq.BatchedEnv(65536, seed=1)
net = xq.DQN(lr=1e-6)
rb = xq.ReplayBuffer(1 << 20)
env.set_stream(s.cuda_stream)
net.set_stream(s.cuda_stream)
xq.collect(net, env, rb, 16, 0.1)
enable_game_events(env, 65536 * 4)
torch.cuda.synchronize()


def timed(f, n=5):
    ts = []
    for _ in range(n):
        torch.cuda.synchronize(); t0 = time.perf_counter(); f(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    return min(ts) * 1e3, sum(ts) / n * 1e3


c = [0]
def upd():
    xq.td_update_replay_n(net, rb, 4096, 5, c[0], 64, True, 1e-6); c[0] += 64
print("collect 4 plies        ms (min, mean)", timed(lambda: xq.collect(net, env, rb, 4, 0.1)))
print("64 TD updates          ms", timed(upd))
print("collect after update   ms", timed(lambda: (upd(), torch.cuda.synchronize(), None)[2] or xq.collect(net, env, rb, 4, 0.1)))
print("drain events           ms", timed(lambda: drain_game_events(env)))
print("target sync            ms", timed(net.update_target_network))
t0 = time.perf_counter(); rep = train(net, env, rb, n_games=60000, plies_per_round=4, updates_per_round=64, batch=4096, lr=1e-6, autosave_games=0); dt = time.perf_counter() - t0
print("train: rounds", rep["plies"] // 4, "ms per round", dt * 1e3 / (rep["plies"] // 4), rep)
