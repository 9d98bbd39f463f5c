"""Generates tests/golden/*.npz from the reference's OWN sources compiled unmodified
(oracle/_ref/libxq_ref.so; recipe oracle/Makefile).  Run in the build container, where
/root/reference exists:   python tests/golden/make_golden.py
The rules fixtures need no GPU.  `--nn-gpu` additionally runs the reference's own CUDA network
(src/dqn.cu compiled unmodified into oracle/_ref/libxq_ref_cuda.so) and must run on a GPU box;
it writes gpurun_out/golden_nn_refcuda.npz, which is then committed as tests/golden/nn_refcuda.npz.
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import loader as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def ref_get(R, h):
    codes = np.zeros(90, np.uint8)
    meta = np.zeros(4, np.int32)
    R.ref_env_get(h, codes, meta)
    return codes, meta


def rules():
    R = O.ref()
    L = O.oracle()
    assert R is not None, "reference build missing"
    h = C.c_void_p(R.ref_env_new())
    buf = np.zeros(256, np.int32)
    # 1. opening KAT
    n = R.ref_env_all_actions(h, 0, buf)
    opening = buf[:2 * n].reshape(n, 2).copy()
    # 2. traced random rollouts: 8 envs x 600 plies, draws = xq_rng(seed, env, ctr) >> 33
    seed, n_envs, plies = 20240607, 8, 600
    traces = np.zeros((n_envs, plies, 6), np.int32)
    finals = np.zeros((n_envs, 94), np.int32)
    pos_codes, pos_meta, pos_counts, pos_lists = [], [], [], []
    for e in range(n_envs):
        R.ref_env_reset(h)
        draws = (O.rng_draws(seed, e, 0, plies) >> np.uint64(33)).astype(np.uint32)
        # replay ply by ply so positions + ordered lists can be harvested on the way
        for p in range(plies):
            codes, meta = ref_get(R, h)
            k = R.ref_env_all_actions(h, int(meta[1]), buf)
            if (p % 9) == e % 9:
                pos_codes.append(codes.copy()); pos_meta.append(meta.copy()); pos_counts.append(k)
                lst = np.full(128, 0xFFFF, np.uint16)
                lst[:k] = (buf[0:2 * k:2] << 7) | buf[1:2 * k:2]
                pos_lists.append(lst)
            t = np.zeros(6, np.int32)
            got = R.ref_env_rollout_random(h, draws[p:p + 1], 1, t.ctypes.data, None)
            assert got == 1
            traces[e, p] = t
        codes, meta = ref_get(R, h)
        finals[e, :90] = codes
        finals[e, 90:] = meta
    # 3. isValidMove / getValidMoves on arbitrary (mostly unreachable) boards
    rng = np.random.default_rng(5)
    m = 256
    arb_codes = np.zeros((m, 90), np.uint8)
    arb_player = rng.integers(0, 2, m).astype(np.uint8)
    arb_q = np.zeros((m, 48, 4), np.int32)
    arb_valid = np.zeros((m, 48), np.uint8)
    arb_counts = np.zeros(m, np.uint8)
    arb_lists = np.full((m, 128), 0xFFFF, np.uint16)
    for i in range(m):
        k = int(rng.integers(2, 36))
        pos = rng.choice(90, k, replace=False)
        arb_codes[i, pos] = rng.integers(1, 15, k)
        meta = np.array([int(rng.integers(0, 199)), int(arb_player[i]), 0, 0], np.int32)
        R.ref_env_set(h, arb_codes[i], meta)
        cnt = R.ref_env_all_actions(h, int(arb_player[i]), buf)
        assert cnt <= 128
        arb_counts[i] = cnt
        arb_lists[i, :cnt] = (buf[0:2 * cnt:2] << 7) | buf[1:2 * cnt:2]
        for j in range(48):
            if j < 40:
                f = int(rng.choice(pos)); t = int(rng.integers(0, 90))
                q = (f // 9, f % 9, t // 9, t % 9)
            else:
                q = tuple(int(v) for v in rng.integers(-1, 11, 4))
            arb_q[i, j] = q
            arb_valid[i, j] = R.ref_env_is_valid_move(h, *q)
    # 4. reward table: evaluateBoard(Red, mc) on boards with assorted material, every mc 0..200
    rew_codes, rew_mc, rew_val = [], [], []
    for i in range(64):
        codes = arb_codes[i]
        R.ref_env_set(h, codes, np.array([0, 0, 0, 0], np.int32))
        for mc in range(0, 201, 1):
            for pl in (0, 1):
                rew_codes.append(i); rew_mc.append((mc, pl)); rew_val.append(R.ref_env_evaluate(h, pl, mc))
    np.savez_compressed(os.path.join(HERE, "rules_ref.npz"), opening=opening, seed=np.uint64(seed), traces=traces, finals=finals,
                        pos_codes=np.array(pos_codes), pos_meta=np.array(pos_meta), pos_counts=np.array(pos_counts, np.uint8),
                        pos_lists=np.array(pos_lists), arb_codes=arb_codes, arb_player=arb_player, arb_q=arb_q,
                        arb_valid=arb_valid, arb_counts=arb_counts, arb_lists=arb_lists,
                        rew_board=np.array(rew_codes, np.int32), rew_mc=np.array(rew_mc, np.int32), rew_val=np.array(rew_val, np.int32))
    print("rules_ref.npz:", len(pos_codes), "positions,", n_envs * plies, "traced plies,", m, "arbitrary boards")


def nn_cases(rng):
    cases = []
    for layers in ([1260, 128, 8100], [12, 8, 30]):
        nw = sum(layers[i] * layers[i + 1] for i in range(len(layers) - 1))
        nb = sum(layers[1:])
        w = rng.uniform(-0.05, 0.05, nw)
        b = rng.uniform(-0.02, 0.02, nb)
        xs = (rng.random((3, layers[0])) < (32.0 / 1260 if layers[0] == 1260 else 0.3)).astype(np.float64)
        ts = rng.uniform(-1, 1, (3, layers[-1]))
        cases.append((np.array(layers, np.int32), w, b, xs, ts))
    return cases


def nn(lib_name, out_path):
    """forward / backpropagate through the reference DQN class; with libxq_ref_cuda.so this is the
    reference's own CUDA code (src/dqn.cu) and needs a GPU."""
    R = C.CDLL(os.path.join(ROOT, "oracle", "_ref", lib_name))
    R.ref_dqn_new.restype = C.c_void_p
    R.ref_last_error.restype = C.c_char_p
    P = C.c_void_p
    rng = np.random.default_rng(11)
    out = {}
    for ci, (layers, w, b, xs, ts) in enumerate(nn_cases(rng)):
        h = R.ref_dqn_new(layers.ctypes.data_as(P), len(layers))
        assert h, R.ref_last_error()
        h = P(h)
        R.ref_dqn_set_params(h, w.ctypes.data_as(P), b.ctypes.data_as(P))
        qs = np.zeros((len(xs), layers[-1]))
        for i, x in enumerate(xs):
            q = np.zeros(layers[-1])
            assert R.ref_dqn_forward(h, x.ctypes.data_as(P), len(x), q.ctypes.data_as(P), len(q)) == len(q)
            qs[i] = q
        for x, t in zip(xs, ts):
            assert R.ref_dqn_backprop(h, x.ctypes.data_as(P), len(x), t.ctypes.data_as(P), len(t), C.c_double(0.001)) == 0
        w2 = np.zeros_like(w); b2 = np.zeros_like(b)
        R.ref_dqn_get_params(h, w2.ctypes.data_as(P), b2.ctypes.data_as(P))
        out.update({f"layers{ci}": layers, f"w{ci}": w, f"b{ci}": b, f"x{ci}": xs, f"t{ci}": ts, f"q{ci}": qs,
                    f"w_after{ci}": w2, f"b_after{ci}": b2})
        R.ref_dqn_free(h)
    np.savez_compressed(out_path, **out)
    print(out_path, "written")


if __name__ == "__main__":
    if "--nn-gpu" in sys.argv:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        nn("libxq_ref_cuda.so", os.path.join(ROOT, "gpurun_out", "golden_nn_refcuda.npz"))
    else:
        O.build()
        rules()
