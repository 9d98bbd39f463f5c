"""DQN: host-side handle of the Q-network (include/xq.h, xq_dqn_*).

Mirrors the reference's DQN class (include/dqn.h:97-110): selectAction, backpropagate, getQValues,
updateTargetNetwork, saveModel/loadModel, train -- plus the batched tensor-core entry points.
"""
import ctypes as C

import numpy as np

from ._lib import ENV_DTYPE, TRANSITION_DTYPE, check, lib, ptr

AS_WRITTEN, CORRECTED = 0, 1
_P = C.c_void_p
_bound = False


def _bind():
    global _bound
    if _bound:
        return lib()
    L = lib()
    L.xq_dqn_create.argtypes = [_P, C.c_int, C.c_double, C.c_double, C.c_int, C.c_uint64, C.c_int, C.POINTER(_P)]
    for f in ("xq_dqn_destroy", "xq_dqn_sync", "xq_dqn_sync_target"):
        getattr(L, f).argtypes = [_P]
    L.xq_dqn_set_stream.argtypes = [_P, _P]
    L.xq_dqn_num_params.argtypes = [_P, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.xq_dqn_set_params.argtypes = [_P, _P, _P]
    L.xq_dqn_get_params.argtypes = [_P, _P, _P]
    L.xq_dqn_forward.argtypes = [_P, _P, C.c_int64, _P]
    L.xq_dqn_backprop.argtypes = [_P, _P, _P, C.c_int64, C.c_double]
    L.xq_dqn_select_action.argtypes = [_P, _P, C.c_double, _P, C.c_int, C.c_uint32, C.c_uint32, C.POINTER(C.c_int)]
    L.xq_dqn_train.argtypes = [_P, _P, C.c_int, C.c_double, _P, C.c_int, C.c_int, C.c_double]
    L.xq_dqn_forward_boards.argtypes = [_P, _P, C.c_int64, _P]
    L.xq_dqn_td_update.argtypes = [_P, _P, C.c_int64, C.c_int, C.c_double, _P]
    L.xq_dqn_td_update_device.argtypes = [_P, _P, C.c_int64, C.c_int, C.c_double, C.c_int]
    L.xq_dqn_grad_buffer.argtypes = [_P, C.POINTER(_P), C.POINTER(C.c_int64)]
    L.xq_dqn_apply_grads.argtypes = [_P, C.c_double]
    L.xq_dqn_dist_export.argtypes = [_P, _P]
    L.xq_dqn_dist_connect.argtypes = [_P, C.c_int, C.c_int, _P]
    L.xq_dqn_dist_allreduce_apply.argtypes = [_P, C.c_double]
    L.xq_dqn_dist_status.argtypes = [_P, C.POINTER(C.c_int)]
    L.xq_dqn_dist_info.argtypes = [_P, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.xq_dqn_dist_allgather.argtypes = [_P, _P, C.c_int64, _P]
    L.xq_dqn_save.argtypes = [_P, C.c_char_p]
    L.xq_dqn_load.argtypes = [_P, C.c_char_p]
    _bound = True
    return L


class DQN:
    def __init__(self, layer_sizes=(1260, 128, 8100), lr=0.001, gamma=0.99, device=0, seed=0, mode=AS_WRITTEN):
        self._L = _bind()
        self.layers = np.ascontiguousarray(layer_sizes, dtype=np.int32)
        self._h = _P()
        check(self._L.xq_dqn_create(ptr(self.layers), len(self.layers), lr, gamma, device, seed, mode, C.byref(self._h)))
        nw, nb = C.c_int64(), C.c_int64()
        check(self._L.xq_dqn_num_params(self._h, C.byref(nw), C.byref(nb)))
        self.n_weights, self.n_biases = nw.value, nb.value
        self.lr, self.gamma = lr, gamma

    def close(self):
        if self._h:
            self._L.xq_dqn_destroy(self._h)
            self._h = _P()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def set_stream(self, cuda_stream):
        check(self._L.xq_dqn_set_stream(self._h, _P(cuda_stream)))

    def sync(self):
        check(self._L.xq_dqn_sync(self._h))

    def set_params(self, w, b):
        w = np.ascontiguousarray(w, dtype=np.float64)
        b = np.ascontiguousarray(b, dtype=np.float64)
        assert w.size == self.n_weights and b.size == self.n_biases
        check(self._L.xq_dqn_set_params(self._h, ptr(w), ptr(b)))

    def get_params(self):
        w = np.empty(self.n_weights, dtype=np.float64)
        b = np.empty(self.n_biases, dtype=np.float64)
        check(self._L.xq_dqn_get_params(self._h, ptr(w), ptr(b)))
        return w, b

    def get_q_values(self, states):
        """DQN::getQValues for one state [in] or a batch [n, in]"""
        x = np.ascontiguousarray(states, dtype=np.float64)
        one = x.ndim == 1
        x = x.reshape(-1, int(self.layers[0]))
        q = np.empty((len(x), int(self.layers[-1])), dtype=np.float64)
        check(self._L.xq_dqn_forward(self._h, ptr(x), len(x), ptr(q)))
        return q[0] if one else q

    def backpropagate(self, states, targets, lr=None):
        x = np.ascontiguousarray(states, dtype=np.float64).reshape(-1, int(self.layers[0]))
        t = np.ascontiguousarray(targets, dtype=np.float64).reshape(-1, int(self.layers[-1]))
        assert len(x) == len(t)
        check(self._L.xq_dqn_backprop(self._h, ptr(x), ptr(t), len(x), self.lr if lr is None else lr))

    def select_action(self, state, eps, actions, coin31, idx31):
        x = np.ascontiguousarray(state, dtype=np.float64)
        a = np.ascontiguousarray(actions, dtype=np.uint16)
        out = C.c_int()
        check(self._L.xq_dqn_select_action(self._h, ptr(x), eps, ptr(a), len(a), coin31, idx31, C.byref(out)))
        return out.value

    def train(self, state, action, reward, next_state, done, use_target_net=True, lr=0.0):
        s = np.ascontiguousarray(state, dtype=np.float64)
        s2 = None if next_state is None else np.ascontiguousarray(next_state, dtype=np.float64)
        check(self._L.xq_dqn_train(self._h, ptr(s), action, reward, ptr(s2), 1 if done else 0, 1 if use_target_net else 0, lr))

    def update_target_network(self):
        check(self._L.xq_dqn_sync_target(self._h))

    def save_model(self, path):
        check(self._L.xq_dqn_save(self._h, str(path).encode()))

    def load_model(self, path):
        check(self._L.xq_dqn_load(self._h, str(path).encode()))

    # ---- batched tensor-core path ({1260,128,8100}) ----
    def forward_boards(self, recs):
        """Q(s) of packed boards through the tcgen05 path: float32 [n, 8100]"""
        recs = np.ascontiguousarray(recs, dtype=ENV_DTYPE)
        q = np.empty((len(recs), int(self.layers[-1])), dtype=np.float32)
        check(self._L.xq_dqn_forward_boards(self._h, ptr(recs), len(recs), ptr(q)))
        return q

    def td_update(self, batch, use_target_net=False, lr=0.0):
        """one batched TD update on host transitions; returns (loss_sum, q_sum, target_sum)"""
        batch = np.ascontiguousarray(batch, dtype=TRANSITION_DTYPE)
        info = np.zeros(4, dtype=np.float32)
        check(self._L.xq_dqn_td_update(self._h, ptr(batch), len(batch), 1 if use_target_net else 0, lr, ptr(info)))
        return info

    def td_update_device(self, batch_ptr, n, use_target_net=False, lr=0.0, apply=True):
        check(self._L.xq_dqn_td_update_device(self._h, _P(batch_ptr), n, 1 if use_target_net else 0, lr, 1 if apply else 0))

    def grad_buffer(self):
        p, n = _P(), C.c_int64()
        check(self._L.xq_dqn_grad_buffer(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def apply_grads(self, lr=0.0):
        check(self._L.xq_dqn_apply_grads(self._h, lr))

    # ---- multi-GPU gradient exchange over peer memory (one process per GPU) ----
    def dist_export(self):
        """64-byte CUDA IPC handle of this rank's gradient exchange buffer"""
        buf = np.zeros(64, dtype=np.uint8)
        check(self._L.xq_dqn_dist_export(self._h, ptr(buf)))
        return buf

    def dist_connect(self, rank, world, handles):
        """handles: uint8 [world, 64], the dist_export() results of all ranks in rank order"""
        hs = np.ascontiguousarray(handles, dtype=np.uint8).reshape(world, 64)
        check(self._L.xq_dqn_dist_connect(self._h, rank, world, ptr(hs)))

    def dist_allreduce_apply(self, lr=0.0):
        """ONE kernel: signal / wait through peer memory, sum the world gradients in rank order over NVLink, apply the SGD step"""
        check(self._L.xq_dqn_dist_allreduce_apply(self._h, lr))

    def dist_timed_out(self):
        """epoch of the first gradient exchange that timed out (0 = none; sticky: later update calls fail)"""
        v = C.c_int()
        check(self._L.xq_dqn_dist_status(self._h, C.byref(v)))
        return bool(v.value)

    def dist_info(self):
        r, w = C.c_int(), C.c_int()
        check(self._L.xq_dqn_dist_info(self._h, C.byref(r), C.byref(w)))
        return r.value, w.value

    def dist_allgather(self, message):
        """host-level all-gather of one small message (<= 64 KB) per rank through the peer-mapped exchange buffer: [world, len] uint8"""
        m = np.ascontiguousarray(message).view(np.uint8).reshape(-1)
        _, world = self.dist_info()
        out = np.zeros((world, m.size), dtype=np.uint8)
        check(self._L.xq_dqn_dist_allgather(self._h, ptr(m), m.size, ptr(out)))
        return out

    def params_digest(self):
        """128-bit digest of the trained parameters' bytes (replica comparison across ranks)"""
        import hashlib
        w, b = self.get_params()
        return np.frombuffer(hashlib.blake2b(w.tobytes() + b.tobytes(), digest_size=16).digest(), dtype=np.uint8).copy()
