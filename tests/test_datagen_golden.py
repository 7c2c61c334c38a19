"""CPU: the oracle's greedy candidates / plane encoder reproduce the reference GreedyDataGenerator's records
(pi, convert_to_train_data rows) on the fixture made by tests/golden/gen_golden_datagen.py."""
import os

import numpy as np

import oracle as orc
from conftest import GOLDEN


def cand_to_pi(masks):
    """uniform over the candidate (checker, destination) pairs at utils.encode_checker_index positions"""
    n = masks.shape[1]
    pi = np.zeros((n, 294))
    for i in range(n):
        idx = [k * 49 + (c >> 3) * 7 + (c & 7) for k in range(6) for c in range(56) if (int(masks[k, i]) >> c) & 1]
        pi[i, idx] = 1.0 / len(idx)
    return pi


def test_oracle_reproduces_reference_generator_records():
    g = dict(np.load(os.path.join(GOLDEN, "datagen_golden.npz")))
    st = g["state"]
    assert np.array_equal(cand_to_pi(orc.greedy_candidates(st)), g["pi"])           # data_generators.py:45-51
    xy = g["has_xy"].astype(bool)
    assert np.array_equal(orc.encode(st)[xy], g["board_x"][xy])                      # utils.py:66 via to_model_input
    # utils.py:65-71: v = reward for player 1, negated every ply == +1 iff the side to move is the eventual winner
    mover_is_p1 = ((st[4] >> np.uint64(48)) & np.uint64(1)) == 0
    want = np.where(mover_is_p1, g["reward"], -g["reward"])
    assert np.array_equal(want[xy], g["v_y"][xy])
