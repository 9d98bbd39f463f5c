"""GPU parity tests (through the C ABI): CUDA env kernels vs the oracle, bit-exact.

Covers BASELINE config 2 (4096 envs random-policy rollouts from the opening: ordered legal list,
chosen action, next board, moveCount, player, scores, done, winner, reward -- every ply, every env),
the golden fixtures generated from the unmodified reference, arbitrary injected boards, edge cases
(rejected moves, terminal boards, masks) and size-independent properties at large N."""
import numpy as np
import pytest

from conftest import harvest_positions, random_boards, recs_from_codes

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def xq():
    import cn_chess_ai_b200 as m
    return m


def oracle_lists(L, recs):
    n = len(recs)
    counts = np.zeros(n, np.uint8)
    acts = np.zeros((n, 128), np.uint16)
    L.xqo_batch_all_actions(recs.ctypes.data, n, counts, acts)
    acts[np.arange(128)[None, :] >= counts[:, None]] = 0xFFFF
    return counts, acts


def same_recs(a, b):
    return a.tobytes() == b.tobytes()


def test_create_is_opening(xq, O):
    env = xq.BatchedEnv(1000, seed=1)
    got = env.get_boards()
    assert same_recs(got, O.new_envs(1000))
    c, a = env.legal_moves()
    assert (c == 44).all()
    assert (a[:, 0] == (0 << 7 | 9)).all() and (a[:, 44:] == 0xFFFF).all()


def test_golden_positions_and_arbitrary(xq, O, golden):
    for codes, meta, counts, lists in (
            (golden["pos_codes"], golden["pos_meta"], golden["pos_counts"], golden["pos_lists"]),
            (golden["arb_codes"], np.stack([np.zeros(256, np.int32), golden["arb_player"].astype(np.int32), np.zeros(256, np.int32),
                                            np.zeros(256, np.int32)], 1), golden["arb_counts"], golden["arb_lists"])):
        recs = recs_from_codes(O, codes, meta)
        env = xq.BatchedEnv(len(recs))
        env.set_boards(recs)
        c, a = env.legal_moves()
        assert (c == counts).all() and (a == lists).all()
    # stand-alone predicate, one query per env, off-board queries included
    recs = recs_from_codes(O, golden["arb_codes"], np.zeros((256, 4), np.int32))
    env = xq.BatchedEnv(256)
    env.set_boards(recs)
    for j in range(golden["arb_q"].shape[1]):
        v = env.is_valid_move(golden["arb_q"][:, j, :])
        assert (v == golden["arb_valid"][:, j]).all()


def test_golden_traces(xq, O, golden):
    tr = golden["traces"]
    n_envs, plies, _ = tr.shape
    env = xq.BatchedEnv(n_envs, seed=int(golden["seed"]))
    stats, out = env.rollout_random(plies, trace=True)
    for e in range(n_envs):
        assert (out["n_legal"][:, e] == tr[e, :, 0]).all()
        assert ((out["action"][:, e] >> 7) == tr[e, :, 1]).all() and ((out["action"][:, e] & 127) == tr[e, :, 2]).all()
        assert (out["reward"][:, e] == tr[e, :, 3]).all()
        assert ((out["flags"][:, e] & 1) == tr[e, :, 4]).all()
        assert (((out["flags"][:, e] >> 1) & 3) == tr[e, :, 5]).all()
    fin = env.get_boards()
    for e in range(n_envs):
        assert (O.codes_of(fin[e]) == golden["finals"][e][:90]).all()
        assert (fin[e]["move_count"], fin[e]["player"], fin[e]["red_score"], fin[e]["black_score"]) == tuple(golden["finals"][e][90:])


def test_config2_4096_envs_stepwise(xq, O, oracle_lib):
    """BASELINE config 2, every ply checked: list -> idx31 % n -> step, 200 plies, auto-reset"""
    n, plies, seed = 4096, 200, 42
    env = xq.BatchedEnv(n, seed=seed)
    ref = O.new_envs(n)
    ids = np.arange(n, dtype=np.uint64)
    for p in range(plies):
        c, a = env.legal_moves()
        c0, a0 = oracle_lists(oracle_lib, ref)
        assert (c == c0).all() and (a == a0).all(), f"ply {p}"
        draws = O.rng_np(seed, ids, ref["ctr"])
        k = ((draws >> np.uint64(33)).astype(np.uint32) % c).astype(np.int64)
        chosen = a[np.arange(n), k]
        rew, done, win, cap, valid = env.step(chosen, auto_reset=True)
        r0 = np.zeros(n, np.int32)
        d0, w0, cp0, v0 = (np.zeros(n, np.uint8) for _ in range(4))
        oracle_lib.xqo_batch_step(ref.ctypes.data, n, chosen, r0, d0, w0, cp0, v0)
        assert (rew == r0).all() and (done == d0).all() and (win == w0).all() and (cap == cp0).all() and (valid == 1).all()
        for i in np.nonzero(d0)[0]:
            oracle_lib.xqo_reset(ref[i:i + 1].ctypes.data)
        assert same_recs(env.get_boards(), ref), f"ply {p}"
    assert ref["ctr"].min() == plies


def test_config2_fused_rollout_matches_oracle(xq, O, oracle_lib):
    """the fused kernel: 4096 envs x 200 plies in one launch == oracle trajectories, ply by ply"""
    n, plies, seed, id0 = 4096, 200, 7, 1000
    env = xq.BatchedEnv(n, seed=seed, env_id0=id0)
    stats, tr = env.rollout_random(plies, trace=True)
    ref = O.new_envs(n)
    tr0 = np.zeros((plies, n), O.TRACE_DTYPE)
    st0 = np.zeros(1, O.STATS_DTYPE)
    oracle_lib.xqo_rollout_random(ref.ctypes.data, n, id0, seed, plies, tr0.ctypes.data, st0.ctypes.data)
    assert tr.tobytes() == tr0.tobytes()
    assert same_recs(env.get_boards(), ref)
    assert stats.tobytes() == st0[0].tobytes()
    # second launch continues the same trajectories (state + counter carried in HBM)
    stats2, _ = env.rollout_random(57)
    oracle_lib.xqo_rollout_random(ref.ctypes.data, n, id0, seed, 57, None, st0.ctypes.data)
    assert same_recs(env.get_boards(), ref) and stats2.tobytes() == st0[0].tobytes()


def test_rollout_from_injected_boards(xq, O, oracle_lib):
    """mid-game, arbitrary and already-terminal boards as starting points; ragged env count"""
    recs = np.concatenate([harvest_positions(O, 300, 7, 23), random_boards(O, 601, seed=2)])
    recs["ctr"] = np.arange(len(recs)) * 3
    n = len(recs)
    assert n % 128 != 0
    env = xq.BatchedEnv(n, seed=5)
    env.set_boards(recs)
    stats, tr = env.rollout_random(64, trace=True)
    ref = recs.copy()
    tr0 = np.zeros((64, n), O.TRACE_DTYPE)
    st0 = np.zeros(1, O.STATS_DTYPE)
    oracle_lib.xqo_rollout_random(ref.ctypes.data, n, 0, 5, 64, tr0.ctypes.data, st0.ctypes.data)
    bad = np.nonzero((tr.view(np.uint64) != tr0.view(np.uint64)).any(0))[0]
    assert len(bad) == 0, f"{len(bad)} envs differ, first {bad[:5]}"
    assert same_recs(env.get_boards(), ref) and stats.tobytes() == st0[0].tobytes()


def test_team_kernel_both_views(xq, O, oracle_lib, monkeypatch):
    """rollout_team_kernel<4> in both forms -- bitboards selected among registers (the default up to 4,736 envs = one CTA per SM) and read through the
    thread's slice of shared memory (above) -- forced in turn on mid-game, arbitrary and finished boards, and at 9,000 envs by default:
    traces, boards and statistics == the oracle's"""
    base = np.concatenate([harvest_positions(O, 300, 7, 23, seed=12), random_boards(O, 600, seed=5)])
    for n, view in ((len(base) - 7, "0"), (len(base) - 7, "1"), (9000, None)):
        recs = np.concatenate([base] * (n // len(base) + 1))[:n].copy()
        recs["ctr"] = np.arange(n) % 700
        if view is not None:
            monkeypatch.setenv("XQ_TEAM_VIEW", view)
        else:
            monkeypatch.delenv("XQ_TEAM_VIEW", raising=False)
        env = xq.BatchedEnv(n, seed=13, env_id0=40)
        env.set_boards(recs)
        stats, tr = env.rollout_random(70, trace=True)
        ref = recs.copy()
        tr0 = np.zeros((70, n), O.TRACE_DTYPE); st0 = np.zeros(1, O.STATS_DTYPE)
        oracle_lib.xqo_rollout_random(ref.ctypes.data, n, 40, 13, 70, tr0.ctypes.data, st0.ctypes.data)
        assert tr.tobytes() == tr0.tobytes() and same_recs(env.get_boards(), ref) and stats.tobytes() == st0[0].tobytes(), (n, view)
    monkeypatch.delenv("XQ_TEAM_VIEW", raising=False)


def test_board_per_thread_kernel_traces(xq, O, oracle_lib):
    """above 9,472 envs the fused rollout is rollout_lane_kernel (one thread per board, the board in registers): every ply of every
    env against the oracle -- from the opening over more than one game, and resumed from mid-game / arbitrary (non-standard piece
    sets go to the generic kernel) / finished boards; ragged env counts"""
    n, plies, seed, id0 = 16411, 230, 21, 5
    env = xq.BatchedEnv(n, seed=seed, env_id0=id0)
    stats, tr = env.rollout_random(plies, trace=True)
    ref = O.new_envs(n)
    tr0 = np.zeros((plies, n), O.TRACE_DTYPE)
    st0 = np.zeros(1, O.STATS_DTYPE)
    oracle_lib.xqo_rollout_random(ref.ctypes.data, n, id0, seed, plies, tr0.ctypes.data, st0.ctypes.data)
    bad = np.nonzero((tr.view(np.uint64) != tr0.view(np.uint64)).any(0))[0]
    assert len(bad) == 0, f"{len(bad)} envs differ, first {bad[:5]}"
    assert same_recs(env.get_boards(), ref) and stats.tobytes() == st0[0].tobytes()
    mid = harvest_positions(O, 400, 9, 23)
    fin = O.new_envs(64)
    fin["move_count"][:32] = 200
    fin["sq"][32:, 0] &= np.uint32(0xFFF0FFFF)          # Red general gone: restarted, never stepped (chessai.cpp:90,96)
    base = np.concatenate([mid, random_boards(O, 900, seed=4), fin])
    recs = np.concatenate([base] * (12500 // len(base) + 1))
    recs["ctr"] = np.arange(len(recs)) % 1000
    n = len(recs)
    assert n > 12288 and n % 128 != 0
    env = xq.BatchedEnv(n, seed=6)
    env.set_boards(recs)
    stats, tr = env.rollout_random(48, trace=True)
    ref = recs.copy()
    tr0 = np.zeros((48, n), O.TRACE_DTYPE)
    st0 = np.zeros(1, O.STATS_DTYPE)
    oracle_lib.xqo_rollout_random(ref.ctypes.data, n, 0, 6, 48, tr0.ctypes.data, st0.ctypes.data)
    bad = np.nonzero((tr.view(np.uint64) != tr0.view(np.uint64)).any(0))[0]
    assert len(bad) == 0, f"{len(bad)} envs differ, first {bad[:5]}"
    assert same_recs(env.get_boards(), ref) and stats.tobytes() == st0[0].tobytes()


def test_api_mode_on_the_device_equals_fused_rollout(xq, O, oracle_lib):
    """the API-mode path without host traffic -- xq_env_legal_moves_device -> xq_env_pick_random_device -> xq_env_step_device, three
    launches per ply -- walks the very trajectories of the fused kernel and of the oracle (same draws: xq_rng(seed, env id, ply counter));
    the lists of the register-resident list kernel against the oracle's on the way, standard and injected boards"""
    n, plies, seed, id0 = 5000, 150, 13, 40
    env = xq.BatchedEnv(n, seed=seed, env_id0=id0)
    for p in range(plies):
        if p in (0, 37, 149):
            counts, acts = env.legal_moves()
            cur = env.get_boards()
            c0 = np.zeros(n, np.uint8); a0 = np.zeros((n, 128), np.uint16)
            oracle_lib.xqo_batch_all_actions(cur.ctypes.data, n, c0, a0)
            a0[np.arange(128)[None, :] >= c0[:, None]] = 0xFFFF
            assert (counts == c0).all() and (acts == a0).all()
        env.api_ply_device()
    ref = O.new_envs(n)
    st0 = np.zeros(1, O.STATS_DTYPE)
    oracle_lib.xqo_rollout_random(ref.ctypes.data, n, id0, seed, plies, None, st0.ctypes.data)
    assert same_recs(env.get_boards(), ref)
    fused = xq.BatchedEnv(n, seed=seed, env_id0=id0)
    fused.rollout_random(plies)
    assert same_recs(env.get_boards(), fused.get_boards())
    # injected boards: arbitrary piece sets go through the generic kernel, standard ones through the register-resident kernel, in one call
    recs = np.concatenate([harvest_positions(O, 300, 7, 23), random_boards(O, 2001, seed=3)])
    env = xq.BatchedEnv(len(recs), seed=1)
    env.set_boards(recs)
    counts, acts = env.legal_moves()
    c0 = np.zeros(len(recs), np.uint8); a0 = np.zeros((len(recs), 128), np.uint16)
    oracle_lib.xqo_batch_all_actions(recs.ctypes.data, len(recs), c0, a0)
    a0[np.arange(128)[None, :] >= c0[:, None]] = 0xFFFF
    assert (counts == c0).all() and (acts == a0).all()


def test_list_kernels_team_and_board_per_thread(xq, O, oracle_lib, monkeypatch):
    """xq_env_legal_moves through BOTH list kernels -- the team of 4 threads per board (legal_moves_team_kernel, the default up to 18,944 envs)
    and one thread per board (legal_moves_lane_kernel) -- forced in turn on the same boards: ordered lists == the oracle's, bit for bit, for
    reachable positions, arbitrary standard boards, arbitrary piece sets (generic kernel) and a grid whose last CTA is partly empty"""
    base = np.concatenate([harvest_positions(O, 700, 9, 33, seed=4), random_boards(O, 1500, seed=17)])
    for n in (len(base) - 13, 3 * len(base) + 5):
        recs = np.concatenate([base] * 4)[:n].copy()
        c0 = np.zeros(n, np.uint8); a0 = np.zeros((n, 128), np.uint16)
        oracle_lib.xqo_batch_all_actions(recs.ctypes.data, n, c0, a0)
        a0[np.arange(128)[None, :] >= c0[:, None]] = 0xFFFF
        env = xq.BatchedEnv(n, seed=1)
        env.set_boards(recs)
        for team in ("1", "0"):
            monkeypatch.setenv("XQ_LEGAL_TEAM", team)
            counts, acts = env.legal_moves()
            assert (counts == c0).all() and (acts == a0).all(), (n, team)
        monkeypatch.delenv("XQ_LEGAL_TEAM")


def test_step_rejects_invalid_moves(xq, O, oracle_lib):
    recs = np.concatenate([harvest_positions(O, 256, 4, 31), random_boards(O, 1024, seed=8)])
    n = len(recs)
    rng = np.random.default_rng(0)
    env = xq.BatchedEnv(n)
    env.set_boards(recs)
    ref = recs.copy()
    for it in range(6):
        acts = ((rng.integers(0, 92, n) << 7) | rng.integers(0, 92, n)).astype(np.uint16)   # mostly illegal, some off-board
        c, a = env.legal_moves()
        pick = rng.random(n) < 0.4
        k = (rng.integers(0, 1 << 30, n) % np.maximum(c, 1)).astype(np.int64)
        acts[pick & (c > 0)] = a[np.arange(n), k][pick & (c > 0)]
        rew, done, win, cap, valid = env.step(acts, auto_reset=False)
        r0 = np.zeros(n, np.int32)
        d0, w0, cp0, v0 = (np.zeros(n, np.uint8) for _ in range(4))
        oracle_lib.xqo_batch_step(ref.ctypes.data, n, acts, r0, d0, w0, cp0, v0)
        assert (valid == v0).all() and (cap == cp0).all() and (rew == r0).all() and (done == d0).all() and (win == w0).all()
        assert same_recs(env.get_boards(), ref)
        assert 0 < valid.sum() < n


def test_valid_moves_per_square_and_state(xq, O, oracle_lib):
    recs = np.concatenate([harvest_positions(O, 128, 3, 40), random_boards(O, 256, seed=12)])
    n = len(recs)
    env = xq.BatchedEnv(n)
    env.set_boards(recs)
    for sq in list(range(0, 90, 5)) + [89]:
        c, to = env.valid_moves(sq // 9, sq % 9)
        for i in range(n):
            t0 = np.zeros(32, np.uint8)
            n0 = oracle_lib.xqo_valid_moves(recs[i:i + 1].ctypes.data, sq // 9, sq % 9, t0)
            assert c[i] == n0 and (to[i, :n0] == t0[:n0]).all()
    c, _ = env.valid_moves(-1, 3)
    assert (c == 0).all()
    st = env.state_onehot()
    for i in range(0, n, 7):
        s0 = np.zeros(1260)
        oracle_lib.xqo_state(recs[i:i + 1].ctypes.data, s0)
        assert (st[i] == s0).all()


def test_reset_mask_and_errors(xq, O):
    env = xq.BatchedEnv(777, seed=3)
    env.rollout_random(33)
    before = env.get_boards()
    mask = (np.arange(777) % 3 == 0).astype(np.uint8)
    env.reset(mask)
    after = env.get_boards()
    fresh = O.new_envs(1)[0]
    for i in range(777):
        if mask[i]:
            assert after[i]["sq"].tobytes() == fresh["sq"].tobytes() and after[i]["move_count"] == 0 and after[i]["ctr"] == before[i]["ctr"]
        else:
            assert after[i].tobytes() == before[i].tobytes()
    with pytest.raises(xq.XQError):
        env.set_boards(before, first=5)
    with pytest.raises(xq.XQError):
        xq.BatchedEnv(0)
    with pytest.raises(xq.XQError):
        xq.BatchedEnv(4, device=99)


def test_large_n_properties(xq, O, oracle_lib):
    """1M envs (BASELINE config 5 size): invariants + an oracle-checked sample of env slots"""
    n, plies, seed = 1 << 20, 64, 11
    env = xq.BatchedEnv(n, seed=seed)
    stats, _ = env.rollout_random(plies)
    assert stats["steps"] == n * plies
    recs = env.get_boards()
    assert (recs["ctr"] == plies).all() and (recs["move_count"] <= 199).all()
    assert (recs["player"] == recs["move_count"] % 2).all()
    assert stats["red_wins"] + stats["black_wins"] == stats["games"]
    idx = np.concatenate([np.arange(0, 300), np.arange(n - 300, n), np.random.default_rng(1).integers(0, n, 400)])
    for i in idx:
        ref = O.new_envs(1)
        st0 = np.zeros(1, O.STATS_DTYPE)
        oracle_lib.xqo_rollout_random(ref.ctypes.data, 1, int(i), seed, plies, None, st0.ctypes.data)
        assert ref[0].tobytes() == recs[int(i)].tobytes()
    # sharding independence: the same global env ids on a differently-offset handle give the same boards
    env2 = xq.BatchedEnv(1000, seed=seed, env_id0=n - 1000)
    env2.rollout_random(plies)
    assert same_recs(env2.get_boards(), recs[n - 1000:])


def test_rollout_random_io_one_call(xq, O, oracle_lib):
    """xq_env_rollout_random_io (host boards in, host boards + trace + stats out, one synchronisation) == set_boards + rollout + get_boards
    == the oracle, resumed from mid-game positions with both colours to move"""
    n, plies, seed = 1500, 60, 19
    start = harvest_positions(O, n // 3, 3, 41, seed=6)
    env = xq.BatchedEnv(n, seed=seed, env_id0=77)
    out = np.empty(n, xq.ENV_DTYPE)
    st, tr = env.rollout_random_io(start, plies, out, trace=True)
    ref = start.copy()
    tr0 = np.zeros((plies, n), O.TRACE_DTYPE)
    st0 = np.zeros(1, O.STATS_DTYPE)
    oracle_lib.xqo_rollout_random(ref.ctypes.data, n, 77, seed, plies, tr0.ctypes.data, st0.ctypes.data)
    assert tr.tobytes() == tr0.tobytes() and same_recs(out, ref) and st.tobytes() == st0[0].tobytes()
    assert same_recs(env.get_boards(), ref)
    st2, _ = env.rollout_random_io(None, 5, None)          # continue on the device, nothing copied back
    oracle_lib.xqo_rollout_random(ref.ctypes.data, n, 77, seed, 5, None, st0.ctypes.data)
    assert same_recs(env.get_boards(), ref) and int(st2["steps"]) == 5 * n
    # PINNED host buffers are read / written by the rollout kernel itself (mapped memory, no copy launches): same results, for the team kernel
    # (1,500 envs) and the board-per-thread kernel (13,000 envs), including boards with non-standard piece sets (generic kernel) and finished ones
    import torch
    for n2 in (1500, 13000):
        base = np.concatenate([harvest_positions(O, 200, 3, 41, seed=6), random_boards(O, 333, seed=9)])
        start2 = np.concatenate([base] * (n2 // len(base) + 1))[:n2].copy()
        start2["ctr"] = np.arange(n2) % 500
        pin_in = torch.from_numpy(start2.view(np.uint8).copy()).pin_memory()
        pin_out = torch.zeros(n2 * 64, dtype=torch.uint8).pin_memory()
        e2 = xq.BatchedEnv(n2, seed=23, env_id0=5)
        st3, tr3 = e2.rollout_random_io(pin_in.numpy().view(xq.ENV_DTYPE), 40, pin_out.numpy().view(xq.ENV_DTYPE), trace=True)
        ref2 = start2.copy()
        tr4 = np.zeros((40, n2), O.TRACE_DTYPE); st4 = np.zeros(1, O.STATS_DTYPE)
        oracle_lib.xqo_rollout_random(ref2.ctypes.data, n2, 5, 23, 40, tr4.ctypes.data, st4.ctypes.data)
        assert tr3.tobytes() == tr4.tobytes() and st3.tobytes() == st4[0].tobytes(), n2
        assert same_recs(pin_out.numpy().view(xq.ENV_DTYPE), ref2) and same_recs(e2.get_boards(), ref2), n2
        assert pin_in.numpy().tobytes() == start2.tobytes()                     # the input buffer is only read


def test_rollout_random_io_submit_wait_double_buffered(xq, O, oracle_lib):
    """xq_env_rollout_random_io_submit / _wait: two env handles alternate on ONE stream, step i + 1 submitted before step i is waited for;
    every step's boards and statistics equal the oracle's, a second submission without a wait is refused (XQ_ERR_STATE)"""
    import torch
    n, plies, rounds = 2048, 25, 4
    s = torch.cuda.Stream()
    envs, pins, refs = [], [], []
    for k in range(2):
        e = xq.BatchedEnv(n, seed=31 + k, env_id0=1000 * k)
        e.set_stream(s.cuda_stream)
        start = harvest_positions(O, n // 4, 3, 41, seed=11 + k)
        start = np.concatenate([start] * (n // len(start) + 1))[:n].copy()
        start["ctr"] = (np.arange(n) * 7 + k) % 900
        pin_in = torch.from_numpy(start.view(np.uint8).copy()).pin_memory()
        pin_out = torch.zeros(n * 64, dtype=torch.uint8).pin_memory()
        pin_st = torch.zeros(64, dtype=torch.uint8).pin_memory()
        envs.append(e); refs.append(start.copy())
        pins.append((pin_in.numpy().view(xq.ENV_DTYPE), pin_out.numpy().view(xq.ENV_DTYPE), pin_st.numpy().view(xq.STATS_DTYPE), (pin_in, pin_out, pin_st)))

    def check_step(k):
        st0 = np.zeros(1, O.STATS_DTYPE)
        oracle_lib.xqo_rollout_random(refs[k].ctypes.data, n, 1000 * k, 31 + k, plies, None, st0.ctypes.data)
        assert same_recs(pins[k][1], refs[k]), k
        assert pins[k][2][0].tobytes() == st0[0].tobytes(), k
        pins[k][0][:] = pins[k][1]                       # the next step of this handle continues from its own output, through the host

    envs[0].rollout_random_io_submit(pins[0][0], plies, pins[0][1], stats_out=pins[0][2])
    with pytest.raises(xq.XQError):
        envs[0].rollout_random_io_submit(pins[0][0], plies, pins[0][1], stats_out=pins[0][2])
    for r in range(rounds):
        envs[1].rollout_random_io_submit(pins[1][0], plies, pins[1][1], stats_out=pins[1][2])
        envs[0].rollout_random_io_wait(); check_step(0)
        if r + 1 < rounds:
            envs[0].rollout_random_io_submit(pins[0][0], plies, pins[0][1], stats_out=pins[0][2])
        envs[1].rollout_random_io_wait(); check_step(1)
    with pytest.raises(xq.XQError):
        envs[1].rollout_random_io_wait()
