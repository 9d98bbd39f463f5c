"""profiling driver: fill a replay ring by self-play, then a few batch-4096 TD updates (used under ncu)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import cn_chess_ai_b200 as xq  # noqa: E402

s = torch.cuda.current_stream()
env = xq.BatchedEnv(65536, seed=1)
net = xq.DQN(lr=1e-6)
rb = xq.ReplayBuffer(1 << 20)
env.set_stream(s.cuda_stream)
net.set_stream(s.cuda_stream)
xq.collect(net, env, rb, 16, 0.1)
for i in range(8):
    xq.td_update_replay(net, rb, 4096, 5, i, True, 1e-6)
torch.cuda.synchronize()
print("ok")
