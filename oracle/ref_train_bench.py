"""TEST / BENCH INFRASTRUCTURE -- times the reference's OWN training loop (ChessAI::train, src/chessai.cpp:85-170: per ply
selectAction -> movePiece -> evaluateBoard -> 2 x getQValues -> backpropagate, batch 1) for a bounded number of games.

Run by bench.py's reference / cpu_baseline legs in a subprocess:  python -m oracle.ref_train_bench [games]
Uses oracle/_ref/libxq_ref_cuda.so (the reference's src/dqn.cu compiled unmodified for sm_100a: its own one-thread-per-neuron
kernels with cudaMalloc / synchronize / cudaFree around every call) when a GPU is present, else oracle/_ref/libxq_ref.so
(same loop, CPU definition of NeuralNetwork from oracle/nn_cpu.cpp).  Prints one JSON line."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def run(libname, games):
    L = C.CDLL(os.path.join(HERE, "_ref", libname))
    L.ref_env_new.restype = C.c_void_p
    L.ref_last_error.restype = C.c_char_p
    L.ref_ai_train.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.ref_env_get.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.ref_rand_load.argtypes = [C.c_void_p, C.c_long]
    h = L.ref_env_new()
    codes = np.zeros(90, np.uint8)
    meta = np.zeros(4, np.int32)
    rng = np.random.default_rng(1)
    plies, secs = 0, 0.0
    for g in range(games + 1):                      # game 0 = warm-up (CUDA context, first allocations)
        draws = rng.integers(0, 2 ** 31 - 1, 1000, dtype=np.int64).astype(np.int32)      # rand() stream of DQN::selectAction (src/dqn.cpp:30-33)
        L.ref_rand_load(draws.ctypes.data, len(draws))
        t0 = time.perf_counter()
        if L.ref_ai_train(h, 1, None, None, None, None) != 0:
            raise RuntimeError(L.ref_last_error().decode())
        dt = time.perf_counter() - t0
        L.ref_env_get(h, codes.ctypes.data, meta.ctypes.data)
        if g > 0:
            plies += int(meta[0]); secs += dt          # ChessBoard::moveCount of the finished game = plies trained on
    return plies, secs


def main():
    games = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    cpu_only = "--cpu" in sys.argv          # the same loop with the network on ONE host core
    out = None
    for lib, kind in (() if cpu_only else (("libxq_ref_cuda.so", "reference ChessAI::train, its own CUDA kernels (src/dqn.cu unmodified, sm_100a) on this GPU"),)) + (("libxq_ref.so", "reference ChessAI::train, NeuralNetwork defined on the CPU (oracle/nn_cpu.cpp), 1 host thread"),):
        if not os.path.exists(os.path.join(HERE, "_ref", lib)):
            continue
        try:
            plies, secs = run(lib, games)
            out = {"transitions_per_s": plies / secs, "plies": plies, "seconds": secs, "games": games, "kind": kind, "batch": 1}
            break
        except Exception as ex:      # no GPU for the CUDA build: fall through to the CPU network
            err = str(ex)
            out = {"unavailable": err}
    print(json.dumps(out if out is not None else {"unavailable": "oracle/_ref is not built"}))


if __name__ == "__main__":
    main()
