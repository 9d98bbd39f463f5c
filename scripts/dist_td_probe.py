"""TD-update timing probe for N ranks (torchrun): K calls x M pipelined updates, per-call us/update (max over ranks) and the CPU time spent
enqueueing.  Knobs: XQ_TD_GEMM_SPLITS, XQ_TD_EARLY_GEMM, XQ_PROBE_MAIN_PRIO=1 (main stream on a higher-priority stream), XQ_DIST_FUSED_MODE."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import cn_chess_ai_b200 as xq
from cn_chess_ai_b200.dist import connect_peers
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
s = torch.cuda.Stream(priority=-1) if os.environ.get("XQ_PROBE_MAIN_PRIO") == "1" else torch.cuda.current_stream()
with torch.cuda.stream(s):
    env = xq.BatchedEnv(65536, device=local, seed=1, env_id0=rank * 65536); net = xq.DQN(lr=1e-6, device=local, seed=3); rb = xq.ReplayBuffer(1 << 20, device=local)
    env.set_stream(s.cuda_stream); net.set_stream(s.cuda_stream)
    xq.collect(net, env, rb, 16, 0.1)
    if world > 1:
        connect_peers(net, dev)
    K, M = int(os.environ.get("XQ_PROBE_CALLS", 10)), int(os.environ.get("XQ_PROBE_UPDATES", 64))
    xq.td_update_replay_n(net, rb, 4096, 5 + rank, 0, 16, True, 1e-6)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    cpu = []
    evs[0].record(s)
    for c in range(K):
        t0 = time.perf_counter()
        xq.td_update_replay_n(net, rb, 4096, 5 + rank, 100 + c * M, M, True, 1e-6)
        cpu.append(1e6 * (time.perf_counter() - t0) / M)
        evs[c + 1].record(s)
    torch.cuda.synchronize()
    per = torch.tensor([1e3 * evs[c].elapsed_time(evs[c + 1]) / M for c in range(K)] + cpu, dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(per, op=dist.ReduceOp.MAX)
    if rank == 0:
        knobs = {k: os.environ[k] for k in ("XQ_TD_GEMM_SPLITS", "XQ_TD_EARLY_GEMM", "XQ_PROBE_MAIN_PRIO", "XQ_DIST_FUSED_MODE", "XQ_TD_GRAPH", "XQ_TD_MAIN_PRIO", "XQ_TD_GEMM_A_TILES", "XQ_TD_FIRST_LATE") if k in os.environ}
        print(f"world {world} {knobs}: us/update per call {[round(float(x), 1) for x in per[:K]]} mean {float(per[:K].mean()):.2f}; cpu enqueue us/update (max over ranks) {[round(float(x), 1) for x in per[K:]]}", flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
