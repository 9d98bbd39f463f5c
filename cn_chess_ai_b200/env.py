"""BatchedEnv: host-side handle of the batched Xiangqi environment (include/xq.h, xq_env_*).

Mirrors, for N boards at once, the board-facing API the reference's training loop uses
(ChessBoard::reset/getValidMoves/movePiece/checkGameOver/getWinner and
ChessAI::getAllValidActions/evaluateBoard/getStateRepresentation).
"""
import ctypes as C

import numpy as np

from ._lib import ENV_DTYPE, MAX_ACTIONS, STATE_SIZE, STATS_DTYPE, TRACE_DTYPE, check, lib, ptr


def action(from_sq, to_sq):
    return (from_sq << 7) | to_sq


class BatchedEnv:
    def __init__(self, n_envs, device=0, seed=0, env_id0=0):
        self._L = lib()
        self._h = C.c_void_p()
        check(self._L.xq_env_create(n_envs, device, seed, env_id0, C.byref(self._h)))
        self.n = int(n_envs)
        self.device = device

    def close(self):
        if self._h:
            self._L.xq_env_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def set_stream(self, cuda_stream):
        check(self._L.xq_env_set_stream(self._h, C.c_void_p(cuda_stream)))

    def sync(self):
        check(self._L.xq_env_sync(self._h))

    def device_boards_ptr(self):
        p = C.c_void_p()
        check(self._L.xq_env_device_boards(self._h, C.byref(p)))
        return p.value

    def reset(self, mask=None):
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        check(self._L.xq_env_reset(self._h, ptr(m)))

    def set_boards(self, recs, first=0):
        recs = np.ascontiguousarray(recs, dtype=ENV_DTYPE)
        check(self._L.xq_env_set_boards(self._h, ptr(recs), first, len(recs)))

    def get_boards(self, first=0, n=None, out=None):
        n = self.n - first if n is None else n
        out = np.empty(n, dtype=ENV_DTYPE) if out is None else out
        check(self._L.xq_env_get_boards(self._h, ptr(out), first, n))
        return out

    def legal_moves(self, strict=False):
        """reference-ordered action lists; strict=True (opt-in, not the reference's rules) drops self-check and flying-general moves"""
        counts = np.empty(self.n, dtype=np.uint8)
        actions = np.empty((self.n, MAX_ACTIONS), dtype=np.uint16)
        check((self._L.xq_env_legal_moves_strict if strict else self._L.xq_env_legal_moves)(self._h, ptr(counts), ptr(actions)))
        return counts, actions

    def valid_moves(self, row, col):
        counts = np.empty(self.n, dtype=np.uint8)
        to = np.empty((self.n, 20), dtype=np.uint8)
        check(self._L.xq_env_valid_moves(self._h, row, col, ptr(counts), ptr(to)))
        return counts, to

    def is_valid_move(self, moves):
        moves = np.ascontiguousarray(moves, dtype=np.int32).reshape(self.n, 4)
        valid = np.empty(self.n, dtype=np.uint8)
        check(self._L.xq_env_is_valid_move(self._h, ptr(moves), ptr(valid)))
        return valid

    def step(self, actions, auto_reset=False):
        actions = np.ascontiguousarray(actions, dtype=np.uint16)
        assert actions.shape == (self.n,)
        reward = np.empty(self.n, dtype=np.int32)
        done, winner, captured, valid = (np.empty(self.n, dtype=np.uint8) for _ in range(4))
        check(self._L.xq_env_step(self._h, ptr(actions), ptr(reward), ptr(done), ptr(winner), ptr(captured), ptr(valid),
                                  1 if auto_reset else 0))
        return reward, done, winner, captured, valid

    def rollout_random(self, n_plies, trace=False):
        tr = np.empty((n_plies, self.n), dtype=TRACE_DTYPE) if trace else None
        stats = np.zeros(1, dtype=STATS_DTYPE)
        check(self._L.xq_env_rollout_random(self._h, n_plies, ptr(tr), ptr(stats)))
        return stats[0], tr

    def rollout_random_io(self, boards_in, n_plies, boards_out, trace=False):
        """host boards in -> n_plies of random-policy self-play -> host boards out, one C-ABI call (one synchronisation)"""
        assert boards_in is None or boards_in.shape == (self.n,)
        assert boards_out is None or boards_out.shape == (self.n,)
        tr = np.empty((n_plies, self.n), dtype=TRACE_DTYPE) if trace else None
        stats = np.zeros(1, dtype=STATS_DTYPE)
        check(self._L.xq_env_rollout_random_io(self._h, ptr(boards_in), n_plies, ptr(boards_out), ptr(tr), ptr(stats)))
        return stats[0], tr

    def rollout_random_io_submit(self, boards_in, n_plies, boards_out, stats_out=None, trace_out=None):
        """first half of rollout_random_io: enqueue and return; the buffers (pinned for the overlap to be real) belong to the library until
        rollout_random_io_wait()"""
        check(self._L.xq_env_rollout_random_io_submit(self._h, ptr(boards_in), n_plies, ptr(boards_out), ptr(trace_out), ptr(stats_out)))

    def rollout_random_io_wait(self):
        check(self._L.xq_env_rollout_random_io_wait(self._h))

    def api_ply_device(self, auto_reset=True):
        """one ply of the API-mode path entirely on the device: ordered lists -> random-policy pick -> movePiece / reward / terminal"""
        check(self._L.xq_env_legal_moves_device(self._h, None, None))
        check(self._L.xq_env_pick_random_device(self._h, None))
        check(self._L.xq_env_step_device(self._h, None, 1 if auto_reset else 0, None, None, None, None, None))

    def legal_moves_device(self):
        check(self._L.xq_env_legal_moves_device(self._h, None, None))

    def rollout_random_async(self, n_plies):
        check(self._L.xq_env_rollout_random_async(self._h, n_plies))

    def stats(self, reset=False):
        stats = np.zeros(1, dtype=STATS_DTYPE)
        check(self._L.xq_env_get_stats(self._h, ptr(stats), 1 if reset else 0))
        return stats[0]

    def state_onehot(self):
        out = np.empty((self.n, STATE_SIZE), dtype=np.float64)
        check(self._L.xq_env_state_onehot(self._h, ptr(out)))
        return out
