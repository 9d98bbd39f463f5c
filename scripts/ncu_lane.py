"""short program for ncu: the board-per-thread rollout at 1M envs, the team kernel at 4096 envs, the list kernel at 65,536 envs"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cn_chess_ai_b200 as xq
s = torch.cuda.current_stream()
big = xq.BatchedEnv(1 << 20, seed=7); big.set_stream(s.cuda_stream)
small = xq.BatchedEnv(4096, seed=7); small.set_stream(s.cuda_stream)
mid = xq.BatchedEnv(65536, seed=7); mid.set_stream(s.cuda_stream)
mid.rollout_random_async(30)
for _ in range(3):
    big.rollout_random_async(32)
    small.rollout_random_async(200)
    mid.api_ply_device()
torch.cuda.synchronize()
print("steps", int(big.stats()["steps"]), int(small.stats()["steps"]))
