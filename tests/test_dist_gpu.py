"""Multi-GPU test of the ONE exchange step of the path: the library's peer-memory gradient exchanges -- inside the gradient
contraction kernel (owner mode: reduce-scatter + all-gather of the row blocks) and as a kernel of its own -- must give bit-identical
parameters on every rank, equal to the NCCL all-reduce + apply baseline at 2 ranks, over several updates (the two slots alternate),
on every update entry point (single, pipelined, n_updates = 1, online-net bootstrap); and the multi-GPU episode driver must report
the game stream of the single-GPU run over the same env range.  Uses every visible GPU (2..8); skipped with fewer than 2
(run with `gpurun --gpus 2`)."""
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
    nets = [xq.DQN(device=local, lr=1e-4) for _ in range(6)]
    for n in nets:
        n.set_params(w, b); n.set_stream(s.cuda_stream)
    env = xq.BatchedEnv(2048, device=local, seed=5, env_id0=rank * 2048)
    env.set_stream(s.cuda_stream)
    rb = xq.ReplayBuffer(1 << 15, device=local)
    xq.collect(nets[0], env, rb, 12, 0.3)                  # rank-local transitions: the ranks' gradients differ
    for k in (0, 2, 3, 4, 5):
        connect_peers(nets[k], dev)
    g1 = grad_tensor(nets[1], dev)

    def same(a, b, what):
        (wa, ba), (wb, bb) = a.get_params(), b.get_params()
        assert wa.tobytes() == wb.tobytes() and ba.tobytes() == bb.tobytes(), (what, float(np.abs(wa - wb).max()))

    for u in range(5):
        xq.td_update_replay(nets[0], rb, 1024, 100 + rank, u, True, 1e-4, apply=False)
        nets[0].dist_allreduce_apply(1e-4)                 # two kernels: push to every rank, then [wait, sum in rank order, SGD]
        xq.td_update_replay(nets[1], rb, 1024, 100 + rank, u, True, 1e-4, apply=False)
        allreduce_sum_(g1)                                 # baseline: NCCL
        nets[1].apply_grads(1e-4)
        xq.td_update_replay(nets[3], rb, 1024, 100 + rank, u, True, 1e-4, apply=True)     # ONE kernel: contraction -> owner exchange -> SGD
    xq.td_update_replay_n(nets[2], rb, 1024, 100 + rank, 0, 5, True, 1e-4)     # the same 5 updates, software-pipelined, exchange per update
    torch.cuda.synchronize()
    assert not any(nets[k].dist_timed_out() for k in (0, 2, 3))
    w0, b0 = nets[0].get_params()
    assert np.abs(w0 - w).max() > 0, "the updates changed nothing"
    if world == 2:      # two addends: every summation order gives the same bits, so the library's exchanges must equal NCCL's sum exactly
        same(nets[0], nets[1], "two-kernel exchange differs from NCCL")
    same(nets[3], nets[0], "exchange inside the contraction (apply=1 on a connected handle) differs from the two-kernel exchange")
    same(nets[2], nets[0], "pipelined multi-GPU updates differ from single updates")
    # the paths the advisor found unexchanged in round 1: n_updates == 1 and the online-net bootstrap (use_target_net = 0)
    xq.td_update_replay_n(nets[4], rb, 1024, 100 + rank, 0, 1, True, 1e-4)
    xq.td_update_replay(nets[5], rb, 1024, 100 + rank, 0, True, 1e-4, apply=True)
    same(nets[4], nets[5], "td_update_replay_n(n_updates=1) on a connected handle")
    for u in range(1, 3):
        xq.td_update_replay_n(nets[4], rb, 1024, 100 + rank, u, 1, False, 1e-4)
        xq.td_update_replay(nets[5], rb, 1024, 100 + rank, u, False, 1e-4, apply=True)
    xq.td_update_replay_n(nets[4], rb, 1024, 100 + rank, 3, 2, False, 1e-4)
    for u in range(3, 5):
        xq.td_update_replay(nets[5], rb, 1024, 100 + rank, u, False, 1e-4, apply=True)
    same(nets[4], nets[5], "online-net bootstrap on a connected handle")
    for k in (0, 2, 3, 4, 5):                              # replicas identical on every rank (digest through the library's own all-gather)
        d = nets[k].dist_allgather(nets[k].params_digest())
        assert d.shape == (world, 16) and (d == d[0]).all(), ("replicas diverged", k)
    msg = np.arange(1000, dtype=np.uint32) * (rank + 1)
    got = nets[0].dist_allgather(msg).view(np.uint32)
    assert all((got[r] == np.arange(1000, dtype=np.uint32) * (r + 1)).all() for r in range(world))

    # config 4 as ONE call per rank: the multi-GPU episode driver.  Same total env range on 1 GPU gives the same game stream.
    def run_train(net_seed, n_envs_rank, id0, connect):
        net = xq.DQN(device=local, lr=1e-5, seed=net_seed); net.set_stream(s.cuda_stream)
        if connect:
            connect_peers(net, dev)
        e = xq.BatchedEnv(n_envs_rank, device=local, seed=9, env_id0=id0); e.set_stream(s.cuda_stream)
        r = xq.ReplayBuffer(1 << 14, device=local)
        seen = []
        rep = xq.train(net, e, r, 400, plies_per_round=40, updates_per_round=0, eps=1.0, train_done=True, autosave_games=0, target_sync_plies=0,
                       on_game_completed=lambda g, a, k: seen.append((g, a, k)))
        return rep, seen, net
    rep, seen, _ = run_train(3, 256, rank * 256, True)     # eps = 1: the trajectories do not depend on the network, so ...
    if rank == 0:
        rep1, seen1, _ = run_train(3, 256 * world, 0, False)      # ... one GPU with the whole env range must report the same games in the same order
        assert seen == seen1 and rep["games"] == rep1["games"] == 400 and rep["transitions"] == rep1["transitions"], (len(seen), len(seen1))
    all_seen = nets[0].dist_allgather(np.asarray(seen[:100], dtype=np.int64))
    assert (all_seen == all_seen[0]).all(), "the ranks report different game streams"
    # ... and with learning: the replicas stay identical through a whole run
    net = xq.DQN(device=local, lr=1e-5, seed=4); net.set_stream(s.cuda_stream)
    connect_peers(net, dev)
    e = xq.BatchedEnv(512, device=local, seed=11, env_id0=rank * 512); e.set_stream(s.cuda_stream)
    r = xq.ReplayBuffer(1 << 15, device=local)
    rep = xq.train(net, e, r, 500, plies_per_round=40, updates_per_round=3, batch=1024, eps=0.1, lr=1e-5, autosave_games=0)
    assert rep["games"] == 500 and rep["updates"] == 3 * rep["plies"] // 40 and rep["transitions"] == rep["plies"] * 512 * world
    d = net.dist_allgather(net.params_digest())
    assert (d == d[0]).all(), "replicas diverged in xq_train_run"
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
    n = min(8, torch.cuda.device_count())
    n = 8 if n >= 8 else (4 if n >= 4 else 2)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
                        "--master-port", "29533", str(script)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "DIST_GPU_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


TIMEOUT_WORKER = textwrap.dedent("""
    import os, sys, time
    import numpy as np
    import torch
    import torch.distributed as dist
    sys.path.insert(0, %r)
    import cn_chess_ai_b200 as xq
    from cn_chess_ai_b200.dist import connect_peers
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    net = xq.DQN(device=local, lr=1e-4, seed=2)
    env = xq.BatchedEnv(1024, device=local, seed=5, env_id0=rank * 1024)
    rb = xq.ReplayBuffer(1 << 14, device=local)
    xq.collect(net, env, rb, 8, 0.3)
    connect_peers(net, dev)
    xq.td_update_replay(net, rb, 512, 1 + rank, 0, True, 1e-4, apply=True)      # a healthy exchange first
    net.sync()
    assert not net.dist_timed_out()
    w_ok, b_ok = net.get_params()
    dist.barrier()
    if rank == 0:       # rank 1 never makes this call: every wait of rank 0's contraction for a peer's row block runs into the time-out
        xq.td_update_replay(net, rb, 512, 1 + rank, 1, True, 1e-4, apply=True)
        net.sync()
        assert net.dist_timed_out(), "the missing peer went unnoticed"
        w1, b1 = net.get_params()
        assert w1.tobytes() == w_ok.tobytes() and b1.tobytes() == b_ok.tobytes(), "a partial sum was applied"
        for call in (lambda: xq.td_update_replay(net, rb, 512, 1, 2, True, 1e-4, apply=True), lambda: xq.td_update_replay_n(net, rb, 512, 1, 2, 3, True, 1e-4),
                     lambda: net.dist_allgather(np.zeros(4, np.uint8))):
            try:
                call()
                raise SystemExit("a call on the failed handle succeeded")
            except xq.XQError as ex:
                assert "timed out" in str(ex), str(ex)
        print("DIST_TIMEOUT_OK")
    else:
        time.sleep(2.0)
    dist.barrier()
    dist.destroy_process_group()
""")


def test_exchange_timeout_is_sticky_and_applies_nothing(tmp_path):
    """a peer that never arrives: the wait inside the contraction kernel ends after XQ_DIST_TIMEOUT_MS, the update is NOT applied (no partial
    sum), and every later update / exchange call on the handle fails with XQ_ERR_STATE (round 1 carried on with stale data after ~2 s)"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "worker_timeout.py"
    script.write_text(TIMEOUT_WORKER % ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29537", str(script)], capture_output=True, text=True, timeout=600, env=dict(os.environ, XQ_DIST_TIMEOUT_MS="300"))
    assert r.returncode == 0 and "DIST_TIMEOUT_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
