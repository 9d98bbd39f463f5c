"""profiling build only (-DXQ_TIMELINE): per-CTA clock64 timeline of the last l1_gemm / dw0_gemm launch, relative to CTA start"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
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
tl = np.zeros((2, 160, 64), np.int64)
L = xq.lib()
L.xq_debug_timeline.argtypes = [C.c_void_p, C.c_int64]
assert L.xq_debug_timeline(tl.ctypes.data, tl.nbytes) == 0
names = {0: {0: "start", 1: "setup done", 2: "W1 tile landed", 3: "end (thread 0)", 36: "end (thread 128)", 37: "end (thread 32)"},
         1: {0: "start", 1: "setup done", 2: "acc_full (epilogue starts)", 3: "tiles parked (cluster barrier 1)", 37: "reduced + applied", 38: "partial written (thread 128)", 39: "idle warp 2 reaches tail", 40: "thread 128 before barrier", 41: "thread 64 before barrier (wold loaded)", 42: "thread 128 after barrier"}}
for k, (kn, ncta) in enumerate((("l1_gemm", 148), ("dw_gemm", 88))):
    t = tl[k, :ncta]
    rel = t - t[:, :1]
    rel[t == 0] = -1
    print(f"== {kn}: cycles since CTA start, mean over {ncta} CTAs [min..max]")
    for slot in range(64):
        v = rel[:, slot]
        v = v[v >= 0]
        if slot and len(v) == 0:
            continue
        if k == 0:
            nm = names[0].get(slot) or (f"MMA: acc_empty ok, tile {slot-4}" if slot < 12 else f"MMA: A landed, tile {slot-12}" if slot < 20 else
                                        f"EPI: acc_full ok, tile {slot-20}" if slot < 28 else f"EPI: drained tile {slot-28}")
        else:
            nm = names[1].get(slot) or (f"MMA: operands ready, kblock {slot-4}" if slot < 12 else f"BLD: words landed, kblock {slot-12}" if slot < 20 else
                                        f"BLD: tile built, kblock {slot-20}")
        print(f"  {slot:2d} {nm:34s} {v.mean():9.0f} [{v.min():7d} .. {v.max():7d}] n={len(v)}")
