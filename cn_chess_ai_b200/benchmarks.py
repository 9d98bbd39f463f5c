"""Timed legs of the DQN half of the metric (imported by bench.py) and the DQN part of smoke()."""
import numpy as np

FLOP_PER_TRANSITION = 9_262_080      # fwd Q(s) + fwd Q'(s') + dense bwd of the {1260,128,8100} MLP


def _td_traffic():
    """DRAM bytes of one TD update from the ncu capture in profiles/ (None if the file is absent)"""
    import json
    import os
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
    return json.load(open(p)).get("td_update_dram_bytes_per_update") if os.path.exists(p) else None


def bench_dqn(stream, peaks, world=1, local=0, dist=None, envs=65536, replay_cap=1 << 20, batch=4096, updates=64, warmup=5):
    """BASELINE configs 3+4 on this rank: eps-greedy self-play with batched Q-net inference fills a 1M-transition
    replay ring, then batch-4096 TD updates (per GPU) are timed with CUDA events on `stream`.
    With world > 1 every update all-reduces the compact gradient over NCCL before the identical SGD step."""
    import torch
    from . import BatchedEnv, DQN, ReplayBuffer, collect, td_update_replay, td_update_replay_n
    import os
    from .dist import allreduce_sum_, connect_peers, grad_tensor
    dev = torch.device("cuda", local)
    env = BatchedEnv(envs, device=local, seed=31, env_id0=local * envs)
    net = DQN((1260, 128, 8100), lr=1e-6, device=local, seed=31)
    rb = ReplayBuffer(replay_cap, device=local)
    env.set_stream(stream.cuda_stream)
    net.set_stream(stream.cuda_stream)
    plies = max(1, replay_cap // envs)
    collect(net, env, rb, 2, 0.1)                      # warm-up
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    collect(net, env, rb, plies, 0.1)
    b.record(stream)
    torch.cuda.synchronize()
    collect_ms = a.elapsed_time(b)
    fused = world > 1 and os.environ.get("XQ_DIST", "fused") != "nccl"
    if fused:
        connect_peers(net, dev)                        # gradient exchange over peer memory, fused with the SGD step
    grads = grad_tensor(net, dev) if world > 1 and not fused else None

    def one(i):
        if fused:
            td_update_replay(net, rb, batch, 1000 + local, i, True, 1e-6, apply=False)
            net.dist_allreduce_apply(1e-6)
        elif world > 1:
            td_update_replay(net, rb, batch, 1000 + local, i, True, 1e-6, apply=False)
            allreduce_sum_(grads)
            net.apply_grads(1e-6)
        else:
            td_update_replay(net, rb, batch, 1000, i, True, 1e-6, apply=True)

    for i in range(warmup):
        one(i)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    pipelined = (world == 1 or fused) and os.environ.get("XQ_TD_PIPELINE", "1") != "0"
    if pipelined:
        td_update_replay_n(net, rb, batch, 1000 + local, 10000, warmup, True, 1e-6)      # warm the second stream / both buffer slots
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
    calls = 3          # timed: `calls` x `updates` sequential updates
    a.record(stream)
    for c in range(calls):
        if pipelined:  # one call = `updates` sequential updates (each with its gradient exchange when world > 1), the target-net branch of
            td_update_replay_n(net, rb, batch, 1000 + local, warmup + c * updates, updates, True, 1e-6)      # update i+1 under the online branch of update i
        else:
            for i in range(updates):
                one(warmup + c * updates + i)
    b.record(stream)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    if fused and net.dist_timed_out():
        raise RuntimeError("gradient exchange: a peer never signalled (dist_timed_out)")
    t = torch.tensor([ms, collect_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, collect_ms = float(t[0]), float(t[1])
    us = 1e3 * ms / (updates * calls)
    tflops = FLOP_PER_TRANSITION * batch / (us * 1e-6) / 1e12          # per GPU
    out = {"metric": "DQN TD updates/s (batch 4096 per GPU, target-net bootstrap, SGD applied)",
           "td_updates_per_s": 1e6 / us, "pipelined_over_two_streams": pipelined, "transitions_per_s": 1e6 / us * batch * world, "us_per_update": us, "batch_per_gpu": batch,
           "replay_transitions_per_gpu": replay_cap,
           "grad_allreduce": ("peer-memory kernel fused with the SGD step" if fused else "nccl all_reduce + apply kernel") if world > 1 else False,
           "selfplay_eps_greedy_steps_per_s": envs * plies * world / (collect_ms * 1e-3), "selfplay_envs_per_gpu": envs,
           "roofline": {"bound": "tensor", "achieved": tflops, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": tflops / peaks["bf16_tflops"],
                        "traffic": _td_traffic(), "peak_source": peaks["source"],
                        "note": "achieved = ALGORITHMIC dense FLOPs (9,262,080 per transition) / time of the whole update (4 kernels chained by programmatic dependent launch, replay draws resolved in place); "
                                "the kernels exploit the one-hot input and one-hot TD error, so far fewer FLOPs are issued (DESIGN.md)"}}
    if world == 1:
        # BASELINE config 4 end to end through the public episode driver (xq_train_run = the batched ChessAI::train): rounds of
        # [4 collector plies over all envs -> 64 batch-4096 TD updates] (replay ratio 1), finished games drained to the host every
        # round; wall clock around the whole call, host work included
        from .trainer import train
        import time
        train(net, env, rb, n_games=2000, plies_per_round=4, updates_per_round=64, batch=batch, lr=1e-6, autosave_games=0)      # warm-up round
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rep = train(net, env, rb, n_games=150000, plies_per_round=4, updates_per_round=64, batch=batch, lr=1e-6, autosave_games=0)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        out["train_loop"] = {"what": "xq_train_run: rounds of 4 eps-greedy plies x %d envs + 64 TD updates of batch %d (target-net bootstrap, target sync every 100 plies), "
                                     "game events drained per round; wall clock" % (envs, batch),
                             "seconds": dt, "rounds": rep["plies"] // 4, "ms_per_round": 1e3 * dt / max(1, rep["plies"] // 4), "games": rep["games"], "games_per_s": rep["games"] / dt, "env_steps_per_s": rep["transitions"] / dt,
                             "td_updates_per_s": rep["updates"] / dt, "trained_transitions_per_s": rep["updates"] * batch / dt}
    env.close(); net.close(); rb.close()
    return out


def smoke_dqn(O):
    """one small forward + TD update of the tensor-core path on cuda:0, checked against the FP64 oracle"""
    from . import DQN, TRANSITION_DTYPE
    L = O.oracle()
    la = np.array([1260, 128, 8100], np.int32)
    rng = np.random.default_rng(0)
    w, b = rng.uniform(-0.05, 0.05, 1260 * 128 + 128 * 8100), rng.uniform(-0.05, 0.05, 128 + 8100)
    net = DQN(la)
    net.set_params(w, b)
    n = 130
    envs = O.new_envs(n)
    st = np.zeros(1, O.STATS_DTYPE)
    for i in range(n):
        L.xqo_rollout_random(envs[i:i + 1].ctypes.data, 1, i, 3, 3 + i, None, st.ctypes.data)
    q = net.forward_boards(envs)
    x = np.zeros((n, 1260)); ref = np.zeros(8100)
    worst = 0.0
    for i in range(n):
        L.xqo_state(envs[i:i + 1].ctypes.data, x[i])
    for i in (0, 64, 129):
        L.xqo_nn_forward(la, 3, w, b, x[i], ref)
        worst = max(worst, float(np.abs(q[i] - ref).max()))
    assert worst < 2e-3, f"tensor-core Q differs from the FP64 oracle by {worst}"
    after = envs.copy()
    counts = np.zeros(n, np.uint8); acts = np.zeros((n, 128), np.uint16)
    L.xqo_batch_all_actions(after.ctypes.data, n, counts, acts)
    pick = acts[np.arange(n), rng.integers(0, 1 << 30, n) % counts]
    rew = np.zeros(n, np.int32); done, win, cap, valid = (np.zeros(n, np.uint8) for _ in range(4))
    L.xqo_batch_step(after.ctypes.data, n, pick, rew, done, win, cap, valid)
    batch = np.zeros(n, TRANSITION_DTYPE)
    batch["s"] = envs["sq"]; batch["s2"] = after["sq"]; batch["action"] = pick; batch["mover"] = envs["player"]
    batch["reward"] = rew; batch["done"] = done
    net.td_update(batch, lr=1e-6)
    gw = np.zeros_like(w); gb = np.zeros_like(b); g1 = np.zeros_like(w); g2 = np.zeros_like(b)
    x2 = np.zeros(1260); qs = np.zeros(8100); qn = np.zeros(8100); tgt = np.zeros(8100)
    for i in range(n):
        L.xqo_state(after[i:i + 1].ctypes.data, x2)
        L.xqo_nn_forward(la, 3, w, b, x[i], qs); L.xqo_nn_forward(la, 3, w, b, x2, qn)
        L.xqo_td_target(qs, qn, 8100, int(pick[i]) & 127, float(rew[i]), int(done[i]), 0.99, tgt)
        L.xqo_nn_grad(la, 3, w, b, x[i], tgt, 0, g1, g2)
        gw += g1; gb += g2
    w1, b1 = net.get_params()
    scale = 1e-6 * max(np.abs(gw).max(), np.abs(gb).max())
    err = max(np.abs((w1 - w) + 1e-6 * gw).max(), np.abs((b1 - b) + 1e-6 * gb).max())
    assert err <= 1e-2 * scale + 1e-7, f"TD update differs from the FP64 oracle: {err} vs scale {scale}"
    net.close()
