"""GPU tests of the episode driver (SURVEY 8f rank 1 + 4): finished-game events of the self-play collector against the oracle,
and xq_train_run's gameCompleted / game_log.txt / autosave / target-sync protocol (src/chessai.cpp:85-170, :370-393)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
LAYERS = [1260, 128, 8100]


@pytest.fixture(scope="module")
def xq():
    import cn_chess_ai_b200 as m
    return m


def test_game_events_match_the_oracle(xq, O, oracle_lib):
    """eps = 1: the collector plays the uniform-random policy, so every finished game (when, which env, scores, move count, winner,
    why) must equal what the oracle's rules produce under the same draws -- in (ply, env) order"""
    n, plies, seed, id0 = 257, 260, 21, 1000                     # > 200 plies: crosses the move cap
    net = xq.DQN(LAYERS, seed=1)
    env = xq.BatchedEnv(n, seed=seed, env_id0=id0)
    xq.enable_game_events(env, n * plies)
    xq.collect(net, env, None, plies, 1.0)
    ev, dropped = xq.drain_game_events(env)
    assert dropped == 0
    ref = O.new_envs(n)
    tr = np.zeros((plies, n), O.TRACE_DTYPE)
    st = np.zeros(1, O.STATS_DTYPE)
    walk = ref.copy()
    oracle_lib.xqo_rollout_random(walk.ctypes.data, n, id0, seed, plies, tr.ctypes.data, st.ctypes.data)
    want = []
    for p in range(plies):
        r0 = np.zeros(n, np.int32); d0, w0, c0, v0 = (np.zeros(n, np.uint8) for _ in range(4))
        before = ref.copy()
        oracle_lib.xqo_batch_step(ref.ctypes.data, n, np.ascontiguousarray(tr[p]["action"]), r0, d0, w0, c0, v0)
        assert (v0 == 1).all()
        for i in np.nonzero(d0)[0]:
            general_gone = int(c0[i]) in (1, 8)                  # captured piece code: a General of either colour
            want.append((p, i, int(ref[i]["red_score"]), int(ref[i]["black_score"]), int(ref[i]["move_count"]), int(w0[i]), 0 if general_gone else 1))
            oracle_lib.xqo_reset(ref[i:i + 1].ctypes.data)
    got = [(int(e["ply"]), int(e["env"]), int(e["red_score"]), int(e["black_score"]), int(e["moves"]), int(e["winner"]), int(e["reason"])) for e in ev]
    assert len(got) == len(want) == int(st[0]["games"]) and len(want) > n
    assert got == want
    assert env.get_boards().tobytes() == ref.tobytes()
    ev2, _ = xq.drain_game_events(env)
    assert len(ev2) == 0                                         # drained


def test_train_run_protocol(xq, tmp_path):
    n_envs, n_games = 512, 700
    log = tmp_path / "game_log.txt"
    prefix = str(tmp_path / "model_after_")

    def run():
        net = xq.DQN(LAYERS, seed=3, lr=1e-5)
        env = xq.BatchedEnv(n_envs, seed=9)
        rb = xq.ReplayBuffer(1 << 16)
        seen = []
        rep = xq.train(net, env, rb, n_games, plies_per_round=40, updates_per_round=2, batch=1024, eps=0.1, lr=1e-5, use_target_net=True,
                       target_sync_plies=100, train_done=True, autosave_games=300, autosave_prefix=prefix, log_path=log, sample_seed=5,
                       on_game_completed=lambda g, r, b: seen.append((g, r, b)))
        w, b = net.get_params()
        return rep, seen, w, b

    rep, seen, w, b = run()
    assert rep["games"] == n_games and [g for g, _, _ in seen] == list(range(1, n_games + 1))       # consecutive game numbers
    assert rep["plies"] % 40 == 0 and rep["transitions"] == rep["plies"] * n_envs and rep["updates"] == 2 * rep["plies"] // 40
    assert rep["target_syncs"] == rep["plies"] // 100 and rep["autosaves"] == 2 and rep["events_dropped"] == 0
    assert rep["red_wins"] == sum(r > k for _, r, k in seen) and rep["black_wins"] == sum(k > r for _, r, k in seen)
    lines = log.read_text().splitlines()
    assert len(lines) == n_games + 2 and lines[-1] == "" and lines[-2] == f"AI self-play session completed. Total games: {n_games}"
    for (g, r, k), line in zip(seen, lines):                     # ChessAI::onGameCompleted's line format
        res = "Red wins!" if r > k else ("Black wins!" if k > r else "It's a draw!")
        assert line == f"Game {g} completed. Red Score: {r}, Black Score: {k}. {res}"
    for g in (300, 600):                                         # DQN::saveModel byte layout: 9,650,484 B for {1260,128,8100}
        assert os.path.getsize(f"{prefix}{g}_games.bin") == 8 * (1260 * 128 + 128 * 8100 + 128 + 8100) + 8 + 4 * 3
    net2 = xq.DQN(LAYERS)
    net2.load_model(f"{prefix}600_games.bin")
    assert np.isfinite(net2.get_params()[0]).all()
    # the whole run is deterministic: same seeds -> same games, same trained weights, bit for bit
    log.unlink()
    rep2, seen2, w2, b2 = run()
    assert seen2 == seen and w2.tobytes() == w.tobytes() and b2.tobytes() == b.tobytes()


def test_self_play_only(xq):
    """ChessAI::startSelfPlay: no learning, done = checkGameOver() only"""
    net = xq.DQN(LAYERS, seed=3)
    env = xq.BatchedEnv(128, seed=2)
    w0, _ = net.get_params()
    seen = []
    rep = xq.train(net, env, None, 150, plies_per_round=64, updates_per_round=0, eps=0.1, train_done=False, autosave_games=0,
                   target_sync_plies=0, on_game_completed=lambda g, r, b: seen.append(g))
    assert rep["games"] == 150 and rep["updates"] == 0 and rep["autosaves"] == 0 and seen == list(range(1, 151))
    assert net.get_params()[0].tobytes() == w0.tobytes()
