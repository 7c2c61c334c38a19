"""chinesecheckersagent_b200 — B200-native batched Chinese Checkers self-play engine.

Python host layer over libccx.so (hand-written sm_100a CUDA behind the C-ABI of include/ccx.h).
The mirror of the reference's Python surface lives in board.py / utils.py / MCTS.py / player.py /
game.py / selfplay.py of this package; the batched API is engine.BatchedEnv.
"""
from . import config  # noqa: F401

__all__ = ["config"]
