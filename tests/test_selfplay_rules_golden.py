"""Pins oracle/selfplay_ref.py (restatement of selfplay.py:29-80's bookkeeping) against logs of the
unmodified selfplay.selfplay() (tests/golden/gen_golden_selfplay.py).  CPU only."""
import os

import numpy as np

import selfplay_ref
from conftest import GOLDEN


def test_replay_matches_reference_selfplay_loop():
    g = np.load(os.path.join(GOLDEN, "selfplay_golden.npz"))
    seen = set()
    for i in range(len(g["n_plies"])):
        n = int(g["n_plies"][i])
        log = g["moves"][i, :n]
        out, _ = selfplay_ref.replay([(int(a), int(b)) for a, b, _, _ in log])
        assert len(out) == n                                   # the game ends exactly where the reference ended it
        for k in range(n):
            assert out[k]["mcts"] == bool(log[k, 2])           # opening length (selfplay.py:32)
            assert out[k]["tau_det"] == bool(log[k, 3])        # tau switch (selfplay.py:62-65)
            assert out[k]["status"] == (0 if k < n - 1 else int(g["status"][i]))
        n_hist = sum(o["mcts"] for o in out) if g["status"][i] in (1, 2) else 0
        assert n_hist == int(g["n_history"][i])
        seen.add(int(g["status"][i]))
    assert seen == {1, 2, 3, 4}
