"""Episode driver: host-side mirror of xq_train_run / xq_env_*_game_events (include/xq.h) -- the batched equivalent of
ChessAI::train / ChessAI::startSelfPlay (src/chessai.cpp:85-170, :191-266) with its gameCompleted signal, game_log.txt
line format, autosave and target-sync cadence."""
import ctypes as C

import numpy as np

from ._lib import check, lib, ptr

_P = C.c_void_p
GAME_EVENT_DTYPE = np.dtype([("ply", "<u4"), ("env", "<u4"), ("red_score", "<i4"), ("black_score", "<i4"), ("moves", "<u2"),
                             ("winner", "u1"), ("reason", "u1"), ("reserved", "<u4")])
assert GAME_EVENT_DTYPE.itemsize == 24
GAME_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_int64, C.c_int32, C.c_int32)


class TrainConfig(C.Structure):
    _fields_ = [("n_games", C.c_int64), ("plies_per_round", C.c_int), ("updates_per_round", C.c_int), ("batch", C.c_int64),
                ("eps", C.c_double), ("lr", C.c_double), ("use_target_net", C.c_int), ("target_sync_plies", C.c_int),
                ("train_done", C.c_int), ("autosave_games", C.c_int), ("autosave_prefix", C.c_char_p), ("log_path", C.c_char_p),
                ("sample_seed", C.c_uint64)]


class TrainReport(C.Structure):
    _fields_ = [(k, C.c_int64) for k in ("games", "plies", "transitions", "updates", "target_syncs", "autosaves", "events_dropped",
                                         "red_wins", "black_wins")] + [("seconds", C.c_double)]


_bound = False


def _bind():
    global _bound
    L = lib()
    if not _bound:
        L.xq_env_enable_game_events.argtypes = [_P, C.c_int64]
        L.xq_env_drain_game_events.argtypes = [_P, _P, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.xq_train_run.argtypes = [_P, _P, _P, C.POINTER(TrainConfig), GAME_CB, _P, C.POINTER(TrainReport)]
        _bound = True
    return L


def enable_game_events(env, capacity):
    check(_bind().xq_env_enable_game_events(env.handle, capacity))
    env._event_cap = int(capacity)


def drain_game_events(env):
    """finished games since the last drain, sorted by (ply, env); returns (events, n_dropped)"""
    out = np.zeros(env._event_cap, dtype=GAME_EVENT_DTYPE)
    n, dropped = C.c_int64(), C.c_int64()
    check(_bind().xq_env_drain_game_events(env.handle, ptr(out), len(out), C.byref(n), C.byref(dropped)))
    return out[:n.value].copy(), dropped.value


def train(net, env, replay, n_games, plies_per_round=16, updates_per_round=4, batch=4096, eps=0.1, lr=0.0, use_target_net=True,
          target_sync_plies=100, train_done=True, autosave_games=100, autosave_prefix=None, log_path=None, sample_seed=0,
          on_game_completed=None):
    """ChessAI::train for env.n boards at once.  on_game_completed(game_number, red_score, blue_score) = the gameCompleted signal.
    updates_per_round = 0, train_done = False gives ChessAI::startSelfPlay without learning.  Returns the report as a dict."""
    L = _bind()
    cfg = TrainConfig(n_games, plies_per_round, updates_per_round, batch, eps, lr, 1 if use_target_net else 0, target_sync_plies,
                      1 if train_done else 0, autosave_games, None if autosave_prefix is None else str(autosave_prefix).encode(),
                      None if log_path is None else str(log_path).encode(), sample_seed)
    rep = TrainReport()
    cb = GAME_CB(lambda user, g, r, b: on_game_completed(int(g), int(r), int(b))) if on_game_completed else C.cast(None, GAME_CB)
    check(L.xq_train_run(net.handle, env.handle, None if replay is None else replay.handle, C.byref(cfg), cb, None, C.byref(rep)))
    return {k: getattr(rep, k) for k, _ in TrainReport._fields_}
