"""Pins the C MCTS oracle against visit counts / pi / Q produced by the unmodified reference MCTS.py
(tests/golden/gen_golden_mcts.py).  Visit counts and Q are bit-exact; pi is bit-exact for tau = 1."""
import os

import numpy as np
import pytest

import oracle as orc
from conftest import GOLDEN


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(os.path.join(GOLDEN, "mcts_golden.npz")))


@pytest.mark.parametrize("ci", range(5))
def test_mcts_oracle_matches_reference(gold, ci):
    evaluator, pre_expand, use_noise, tau, num_itr = gold["configs"][ci]
    v, pi, q, nodes = orc.mcts(gold["roots"], int(num_itr), 3.5, float(tau), int(pre_expand), int(evaluator),
                               gold["noise"] if use_noise else None, nthreads=8)
    assert np.array_equal(v, gold["visits%d" % ci])
    assert np.array_equal(q, gold["q%d" % ci])                 # float64, bit for bit
    assert np.array_equal(nodes, gold["nodes%d" % ci])
    if tau == 1.0:
        assert np.array_equal(pi, gold["pi%d" % ci])
    else:
        assert np.allclose(pi, gold["pi%d" % ci], rtol=1e-12, atol=1e-300)
    assert np.all(v.sum(1) == (175 if pre_expand else 174))


def test_terminal_backups_are_exercised(gold):
    # late greedy roots reach wins inside the horizon: Q = +-1 appears at the root for the hash evaluator
    q = gold["q2"][40:]
    assert np.any(np.abs(q) == 1.0)


@pytest.mark.parametrize("pre_expand", [0, 1])
def test_oracle_matches_reference_on_256_roots(pre_expand):
    """SURVEY §8d cfg 4: >= 256 reference-run roots (tests/golden/gen_golden_mcts256.py), incl. Board(randomised=True) ones."""
    g = dict(np.load(os.path.join(GOLDEN, "mcts_golden_256.npz")))
    assert g["roots"].shape[1] >= 256
    v, pi, q, nodes = orc.mcts(g["roots"], 175, 3.5, 1.0, pre_expand, 0, nthreads=8)
    assert np.array_equal(v, g["visits%d" % pre_expand]) and np.array_equal(q, g["q%d" % pre_expand])
    assert np.array_equal(pi, g["pi%d" % pre_expand]) and np.array_equal(nodes, g["nodes%d" % pre_expand])
