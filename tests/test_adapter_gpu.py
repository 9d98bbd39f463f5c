"""GPU test of the C++ adapter classes (cn_chess_ai_b200/adapter/xq_adapter.hpp): the reference's ChessBoard /
ChessAI / DQN / Action API surface, driven the way MainWindow/Worker drive the originals."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_adapter_transcript(tmp_path, O):
    exe = tmp_path / "test_adapter"
    pkg = os.path.join(ROOT, "cn_chess_ai_b200")
    subprocess.run(["g++", "-std=c++17", "-O1", "-o", str(exe), os.path.join(ROOT, "tests", "cpp", "test_adapter.cpp"), "-L" + pkg,
                    "-lxq_b200", "-Wl,-rpath," + pkg], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600, cwd=str(tmp_path))
    assert out.returncode == 0, out.stderr[-2000:]
    got = dict((ln.split()[0], ln.split()[1:]) for ln in out.stdout.strip().splitlines())
    # known answers of the reference (re-derived from the unmodified build when it is present)
    want = {"opening_actions": ["44", "first", "0", "9"], "valid_2_1_9_1": ["1", "valid_0_0_5_5", "0"],
            "captured": ["4", "1", "red_score", "40", "eval", "39", "-40", "37"],
            "moves_9_0": ["3", "player", "1", "movecount", "1", "over", "0"], "copy_movecount": ["0", "orig_movecount", "1"],
            "state": ["1260", "ones", "31"], "q_size": ["8100", "q_in_range", "1"], "trained_games": ["1", "red_score_nonneg", "1", "dqn", "1"],
            "ai_move_valid": ["1"], "batched_games": ["300", "last", "300", "finished", "1", "report", "300", "updates_positive", "1"]}
    for k, v in want.items():
        assert got[k] == v, (k, got[k], v)
    assert out.stdout.count("invalid_argument") == 1 and out.stdout.count("runtime_error") == 1
    log = (tmp_path / "game_log.txt").read_text().splitlines()
    assert len(log) == 302 and log[0].startswith("Game 1 completed. Red Score: ") and log[300] == "AI self-play session completed. Total games: 300"
    R = O.ref()
    if R is not None:   # the same calls through the reference's own classes
        h = C.c_void_p(R.ref_env_new())
        buf = np.zeros(256, np.int32)
        assert R.ref_env_all_actions(h, 0, buf) == 44 and R.ref_env_is_valid_move(h, 2, 1, 9, 1) == 1
        assert R.ref_env_move(h, 19, 82) == 11 and R.ref_env_evaluate(h, 0, 1) == 39 and R.ref_env_evaluate(h, 1, 1) == -40
        assert R.ref_env_valid_moves(h, 9, 0, buf) == 3
