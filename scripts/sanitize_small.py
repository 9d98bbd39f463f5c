"""small invocation of every kernel family for compute-sanitizer memcheck (one tool, small sizes)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cn_chess_ai_b200 as xq
env = xq.BatchedEnv(12400, seed=3)                 # lane rollout kernel (> 12,288 envs), ragged tail
env.rollout_random(6, trace=True)
env.api_ply_device(); env.api_ply_device()
c, a = env.legal_moves()
small = xq.BatchedEnv(777, seed=4)                 # team rollout kernel
small.rollout_random(9, trace=True)
net = xq.DQN(lr=1e-6, seed=1)
rb = xq.ReplayBuffer(1 << 14)
e2 = xq.BatchedEnv(1500, seed=5)
xq.collect(net, e2, rb, 5, 0.2)
acts = xq.act(net, e2, 0.1)
xq.td_update_replay(net, rb, 1024, 7, 0, True, 1e-6)
xq.td_update_replay_n(net, rb, 1024, 7, 1, 3, True, 1e-6)
net.sync()
w, b = net.get_params()
print("sanitize run ok", int(c.sum()), int(acts[0]), float(np.abs(w).max()))
