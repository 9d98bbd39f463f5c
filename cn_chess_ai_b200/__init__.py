"""cn_chess_ai_b200: B200-native batched Xiangqi environment + DQN trainer (sm_100a only).

Host-side mirror of the C ABI in include/xq.h; the hot path lives in csrc/ as hand-written CUDA.
"""
from ._lib import (ENV_DTYPE, MAX_ACTIONS, STATE_SIZE, STATS_DTYPE, TRACE_DTYPE, TRANSITION_DTYPE, XQError,  # noqa: F401
                   lib)
from .env import BatchedEnv, action  # noqa: F401
from .dqn import AS_WRITTEN, CORRECTED, DQN  # noqa: F401
from .replay import ReplayBuffer, act, collect, td_update_replay, td_update_replay_n  # noqa: F401
from .trainer import GAME_EVENT_DTYPE, drain_game_events, enable_game_events, train  # noqa: F401
from .benchmarks import bench_dqn  # noqa: F401
