"""ReplayBuffer + self-play collection: host-side handles of xq_replay_* / xq_selfplay_collect (include/xq.h)."""
import ctypes as C

import numpy as np

from ._lib import TRANSITION_DTYPE, check, lib, ptr

_P = C.c_void_p
_bound = False


def _bind():
    global _bound
    L = lib()
    if _bound:
        return L
    L.xq_replay_create.argtypes = [C.c_int64, C.c_int, C.POINTER(_P)]
    L.xq_replay_destroy.argtypes = [_P]
    L.xq_replay_info.argtypes = [_P, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.xq_replay_insert.argtypes = [_P, _P, C.c_int64]
    L.xq_replay_get.argtypes = [_P, C.c_int64, C.c_int64, _P]
    L.xq_replay_sample.argtypes = [_P, C.c_int64, C.c_uint64, C.c_uint32, _P, _P]
    L.xq_dqn_act.argtypes = [_P, _P, C.c_double, _P, _P]
    L.xq_selfplay_collect.argtypes = [_P, _P, _P, C.c_int, C.c_double, C.c_int]
    L.xq_dqn_td_update_replay.argtypes = [_P, _P, C.c_int64, C.c_uint64, C.c_uint32, C.c_int, C.c_double, C.c_int]
    L.xq_dqn_td_update_replay_n.argtypes = [_P, _P, C.c_int64, C.c_uint64, C.c_uint32, C.c_int, C.c_int, C.c_double]
    _bound = True
    return L


class ReplayBuffer:
    def __init__(self, capacity, device=0):
        self._L = _bind()
        self._h = _P()
        check(self._L.xq_replay_create(capacity, device, C.byref(self._h)))
        self.capacity = int(capacity)

    def close(self):
        if self._h:
            self._L.xq_replay_destroy(self._h)
            self._h = _P()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def info(self):
        size, cap, total = C.c_int64(), C.c_int64(), C.c_int64()
        check(self._L.xq_replay_info(self._h, C.byref(size), C.byref(cap), C.byref(total)))
        return size.value, cap.value, total.value

    def __len__(self):
        return self.info()[0]

    def insert(self, batch):
        batch = np.ascontiguousarray(batch, dtype=TRANSITION_DTYPE)
        check(self._L.xq_replay_insert(self._h, ptr(batch), len(batch)))

    def get(self, first=0, n=None):
        n = self.capacity - first if n is None else n
        out = np.empty(n, dtype=TRANSITION_DTYPE)
        check(self._L.xq_replay_get(self._h, first, n, ptr(out)))
        return out

    def sample(self, batch, seed, counter):
        out = np.empty(batch, dtype=TRANSITION_DTYPE)
        idx = np.empty(batch, dtype=np.int64)
        check(self._L.xq_replay_sample(self._h, batch, seed, counter, ptr(out), ptr(idx)))
        return out, idx


def act(dqn, env, eps=0.1, want_q=False):
    """DQN::selectAction for every env (nothing applied): actions [n] (+ Q(s)[0..89] as float32 [n, 96])"""
    L = _bind()
    actions = np.empty(env.n, dtype=np.uint16)
    q = np.empty((env.n, 96), dtype=np.float32) if want_q else None
    check(L.xq_dqn_act(dqn.handle, env.handle, eps, ptr(actions), ptr(q)))
    return (actions, q) if want_q else actions


def collect(dqn, env, replay, n_plies, eps=0.1, train_done=True):
    """n_plies of epsilon-greedy self-play on the device; transitions go to `replay` (may be None)"""
    L = _bind()
    check(L.xq_selfplay_collect(dqn.handle, env.handle, replay.handle if replay is not None else None, n_plies, eps,
                                1 if train_done else 0))


def td_update_replay(dqn, replay, batch, seed, counter, use_target_net=False, lr=0.0, apply=True):
    L = _bind()
    check(L.xq_dqn_td_update_replay(dqn.handle, replay.handle, batch, seed, counter, 1 if use_target_net else 0, lr, 1 if apply else 0))


def td_update_replay_n(dqn, replay, batch, seed, counter0, n_updates, use_target_net=True, lr=0.0):
    """n_updates consecutive TD updates (counters counter0, counter0 + 1, ...), software-pipelined over two streams when the
    bootstrap comes from the target net; bit-identical to n_updates calls of td_update_replay"""
    L = _bind()
    check(L.xq_dqn_td_update_replay_n(dqn.handle, replay.handle, batch, seed, counter0, n_updates, 1 if use_target_net else 0, lr))
