"""profiling build only (XQ_LIB_PATH=.../libxq_b200_tl.so, torchrun, one rank per GPU): where the time of the gradient exchange inside
dw_gemm_kernel goes -- per rank, clock64 stamps of the LAST update of a pipelined run, per CTA, in us relative to the CTA's own
"reduced rows ready" stamp: non-owner blocks [sent -> owner's sum in], owner blocks [sources 0..3 in, sources 4..7 in, sum sent]."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import cn_chess_ai_b200 as xq
from cn_chess_ai_b200.dist import connect_peers
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
s = torch.cuda.current_stream()
env = xq.BatchedEnv(65536, device=local, seed=1, env_id0=rank * 65536); net = xq.DQN(lr=1e-6, device=local, seed=3); rb = xq.ReplayBuffer(1 << 20, device=local)
env.set_stream(s.cuda_stream); net.set_stream(s.cuda_stream)
xq.collect(net, env, rb, 16, 0.1)
connect_peers(net, dev)
xq.td_update_replay_n(net, rb, 4096, 5 + rank, 0, 16, True, 1e-6)
dist.barrier(); torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
N = 64
a.record(s)
xq.td_update_replay_n(net, rb, 4096, 5 + rank, 100, N, True, 1e-6)
b.record(s)
torch.cuda.synchronize()
us = 1e3 * a.elapsed_time(b) / N
tl = np.zeros((2, 160, 64), np.int64)
L = xq.lib()
L.xq_debug_timeline.argtypes = [C.c_void_p, C.c_int64]
assert L.xq_debug_timeline(tl.ctypes.data, tl.nbytes) == 0
t = tl[1, :88].astype(np.float64)
mhz = 1965.0
def rel(slot, base=41, sel=None):
    m = (t[:, slot] > 0) & (t[:, base] > 0)
    if sel is not None:
        m &= sel
    v = (t[m, slot] - t[m, base]) / mhz
    return (f"{v.mean():6.2f} [{v.min():6.2f} .. {v.max():6.2f}] n={len(v)}" if len(v) else "n=0")
own = t[:, 45] == 1
start = t[:, 0]
lines = [f"rank {rank}/{world}: {us:.2f} us per update (pipelined, {N} updates)",
         f"  CTA start spread over the 88 CTAs                   {(start.max() - start.min()) / mhz:6.2f} us",
         f"  rows reduced (exchange start) after CTA start        {rel(41, 0)}",
         f"  exchange start spread over the 88 CTAs               {(t[:, 41][t[:, 41] > 0].max() - t[:, 41][t[:, 41] > 0].min()) / mhz:6.2f} us",
         f"  non-owner: copy sent                                 {rel(42, 41, ~own)}",
         f"  non-owner: owner's sum in                            {rel(44, 41, ~own)}",
         f"  owner: sources 0..3 in                               {rel(48, 41, own)}",
         f"  owner: sources 4..7 in                               {rel(49, 41, own)}",
         f"  owner: sum sent                                      {rel(43, 41, own)}",
         f"  end of the CTA (thread 32) after exchange start      {rel(37, 41)}",
         f"  end of the CTA after CTA start                       {rel(37, 0)}"]
for r in range(world):
    dist.barrier()
    if r == rank:
        print("\n".join(lines), flush=True)
dist.barrier()
dist.destroy_process_group()
