"""profiling build only (-DXQ_TIMELINE): cycles per ply that each warp (= piece slot) of block 0 of rollout_slots_kernel spends
in the phases A-D and waiting at the three barriers"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import cn_chess_ai_b200 as xq  # noqa: E402

plies = 200
env = xq.BatchedEnv(4096, seed=2024)
env.rollout_random_async(plies)
env.rollout_random_async(plies)
env.sync()
ph = np.zeros((16, 8), np.int64)
L = xq.lib()
L.xq_debug_rollout_phases.argtypes = [C.c_void_p]
assert L.xq_debug_rollout_phases(ph.ctypes.data) == 0
names = ["A count", "wait 1", "B prefix+decode", "wait 2", "C apply", "wait 3", "D score/out", "-"]
types = ["Chariot"] * 2 + ["Horse"] * 2 + ["Elephant"] * 2 + ["Advisor"] * 2 + ["General"] + ["Cannon"] * 2 + ["Soldier"] * 5
print("cycles per ply, block 0 (32 boards), 4096 envs x %d plies" % plies)
print("slot type     " + "".join(f"{n:>17s}" for n in names[:7]) + "   total")
for s in range(16):
    v = ph[s, :7] / plies
    print(f"{s:2d}   {types[s]:9s}" + "".join(f"{x:17.0f}" for x in v) + f"{v.sum():8.0f}")
