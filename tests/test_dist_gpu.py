"""2-GPU test of the ONE exchange step of the path: the library's peer-memory gradient exchange fused with the
SGD step must give bit-identical parameters to the NCCL all-reduce + apply baseline, on every rank, over several updates
(the two gradient slots alternate).  Skipped on boxes with fewer than 2 GPUs (run with `gpurun --gpus 2`)."""
import os
import subprocess
import sys
import textwrap

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np
    import torch
    import torch.distributed as dist
    sys.path.insert(0, %r)
    import cn_chess_ai_b200 as xq
    from cn_chess_ai_b200.dist import connect_peers, grad_tensor, allreduce_sum_
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    s = torch.cuda.current_stream()
    rng = np.random.default_rng(3)
    w, b = rng.uniform(-0.05, 0.05, 1260 * 128 + 128 * 8100), rng.uniform(-0.05, 0.05, 128 + 8100)
    nets = [xq.DQN(device=local, lr=1e-4) for _ in range(3)]
    for n in nets:
        n.set_params(w, b); n.set_stream(s.cuda_stream)
    env = xq.BatchedEnv(2048, device=local, seed=5, env_id0=rank * 2048)
    env.set_stream(s.cuda_stream)
    rb = xq.ReplayBuffer(1 << 15, device=local)
    xq.collect(nets[0], env, rb, 12, 0.3)                  # rank-local transitions: the ranks' gradients differ
    connect_peers(nets[0], dev)
    connect_peers(nets[2], dev)
    g1 = grad_tensor(nets[1], dev)
    for u in range(5):
        xq.td_update_replay(nets[0], rb, 1024, 100 + rank, u, True, 1e-4, apply=False)
        nets[0].dist_allreduce_apply(1e-4)                 # fused: flags + peer loads over NVLink + SGD in one kernel
        xq.td_update_replay(nets[1], rb, 1024, 100 + rank, u, True, 1e-4, apply=False)
        allreduce_sum_(g1)                                 # baseline: NCCL
        nets[1].apply_grads(1e-4)
    xq.td_update_replay_n(nets[2], rb, 1024, 100 + rank, 0, 5, True, 1e-4)     # the same 5 updates, software-pipelined, exchange per update
    torch.cuda.synchronize()
    assert not nets[0].dist_timed_out() and not nets[2].dist_timed_out()
    w2, b2 = nets[2].get_params()
    w0, b0 = nets[0].get_params(); w1, b1 = nets[1].get_params()
    assert np.abs(w0 - w).max() > 0, "the updates changed nothing"
    assert w0.tobytes() == w1.tobytes() and b0.tobytes() == b1.tobytes(), ("fused exchange differs from NCCL", float(np.abs(w0 - w1).max()))
    assert w2.tobytes() == w0.tobytes() and b2.tobytes() == b0.tobytes(), "pipelined multi-GPU updates differ from single updates"
    digest = torch.tensor([float(np.sum(w0 * np.arange(1, w0.size + 1) %% 977)), float(b0.sum())], dtype=torch.float64, device=dev)
    all_d = [torch.empty_like(digest) for _ in range(world)]
    dist.all_gather(all_d, digest)
    assert all(torch.equal(all_d[0], d) for d in all_d), "replicas diverged"
    dist.barrier()
    if rank == 0:
        print("DIST_GPU_OK")
    dist.destroy_process_group()
""")


def test_peer_memory_exchange_matches_nccl(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", str(script)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "DIST_GPU_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
