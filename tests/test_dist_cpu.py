"""CPU suite: the N>1 path on the gloo backend (world_size 2).  No GPU: what is exercised is the host-side
sharding (env ids, replay capacity) and the one exchange step of the data-parallel learner -- a sum
all-reduce of per-rank gradients must equal the single-rank gradient of the concatenated batch (the
per-rank gradients come from the FP64 oracle here; on the GPU box they come from the TD-update kernels)."""
import os
import subprocess
import sys
import textwrap

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_partition():
    from cn_chess_ai_b200.dist import shard
    for total in (1, 7, 4096, 1 << 20, 1000003):
        for world in (1, 2, 3, 4, 8):
            parts = [shard(total, r, world) for r in range(world)]
            assert parts[0][0] == 0 and sum(c for _, c in parts) == total
            for (f0, c0), (f1, _) in zip(parts, parts[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1


WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np
    import torch
    import torch.distributed as dist
    sys.path.insert(0, %r)
    from oracle import loader as O
    from cn_chess_ai_b200.dist import shard, allreduce_sum_
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    L = O.oracle()
    layers = np.array([1260, 16, 96], np.int32)            # small net, same structure
    rng = np.random.default_rng(0)
    w = rng.uniform(-0.05, 0.05, 1260 * 16 + 16 * 96); b = rng.uniform(-0.05, 0.05, 16 + 96)
    # global batch of 24 positions from oracle rollouts; env ids are GLOBAL, so every rank can rebuild any of them
    total = 24
    def sample(i):
        e = O.new_envs(1); st = np.zeros(1, O.STATS_DTYPE)
        L.xqo_rollout_random(e.ctypes.data, 1, i, 77, 10 + i, None, st.ctypes.data)
        x = np.zeros(1260); L.xqo_state(e.ctypes.data, x)
        t = np.zeros(96); L.xqo_nn_forward(layers, 3, w, b, x, t); t[i %% 90] = float(i) - 10.0
        return x, t
    def grads(ids):
        gw = np.zeros_like(w); gb = np.zeros_like(b); a = np.zeros_like(w); c = np.zeros_like(b)
        for i in ids:
            x, t = sample(i); L.xqo_nn_grad(layers, 3, w, b, x, t, 1, a, c); gw += a; gb += c
        return np.concatenate([gw, gb])
    first, count = shard(total, rank, world)
    local = torch.from_numpy(grads(range(first, first + count)))
    allreduce_sum_(local)
    if rank == 0:
        full = grads(range(total))
        err = float(np.abs(local.numpy() - full).max())
        scale = float(np.abs(full).max())
        assert err <= 1e-12 * max(scale, 1.0), (err, scale)
        print("OK", world, err, scale)
    dist.barrier(); dist.destroy_process_group()
""")


def test_gloo_world2_gradient_allreduce(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29611", str(script)], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "OK 2" in r.stdout
