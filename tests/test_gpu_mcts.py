"""GPU parity tests for the batched MCTS kernels, through the C-ABI.  Visit counts and root Q are
bit-exact against fixtures produced by the unmodified reference MCTS.py and against the C oracle."""
import os

import numpy as np
import pytest
import torch

import oracle as orc
from conftest import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from chinesecheckersagent_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(os.path.join(GOLDEN, "mcts_golden.npz")))


def dev(a):
    a = np.ascontiguousarray(a)
    if a.dtype == np.uint64:
        a = a.view(np.int64)
    return torch.from_numpy(a).cuda()


@pytest.mark.parametrize("ci", range(5))
def test_search_matches_reference_mcts(eng, gold, ci):
    from chinesecheckersagent_b200.engine import BatchedMCTS
    evaluator, pre_expand, use_noise, tau, num_itr = gold["configs"][ci]
    m = BatchedMCTS(eng, cpuct=3.5, num_itr=int(num_itr), tree_tau=float(tau))
    out = m.search(dev(gold["roots"]), evaluator=int(evaluator), pre_expand=bool(pre_expand),
                   root_noise=dev(gold["noise"]) if use_noise else None)
    visits = out["visits"].cpu().numpy().astype(np.uint32)
    assert np.array_equal(visits, gold["visits%d" % ci])
    assert np.array_equal(out["q"].cpu().numpy(), gold["q%d" % ci])           # float64 bit for bit
    assert np.array_equal(out["n_nodes"].cpu().numpy(), gold["nodes%d" % ci])
    pi = out["pi"].cpu().numpy()
    if tau == 1.0:
        assert np.array_equal(pi, gold["pi%d" % ci])
    else:
        assert np.allclose(pi, gold["pi%d" % ci], rtol=1e-9, atol=1e-300)


def cfg4_roots(n, seed=0x5EED2026):
    st, _, _ = orc.step_random(orc.start_states(n), seed, 0, 6, nthreads=8)      # start advanced by 6 random plies
    return st


@pytest.mark.parametrize("evaluator,pre_expand", [(0, 0), (1, 0), (1, 1)])
def test_search_256_roots_vs_oracle(eng, evaluator, pre_expand):
    from chinesecheckersagent_b200.engine import BatchedMCTS
    roots = cfg4_roots(256)
    out = BatchedMCTS(eng).search(dev(roots), evaluator=evaluator, pre_expand=bool(pre_expand))
    v, pi, q, nodes = orc.mcts(roots, 175, 3.5, 1.0, pre_expand, evaluator, nthreads=8)
    assert np.array_equal(out["visits"].cpu().numpy().astype(np.uint32), v)
    assert np.array_equal(out["q"].cpu().numpy(), q)
    assert np.array_equal(out["pi"].cpu().numpy(), pi)
    assert np.array_equal(out["n_nodes"].cpu().numpy(), nodes)


def test_round_based_search_equals_persistent(eng):
    """The select / evaluate / expand+backup pipeline used for the net gives the same trees as the
    persistent kernel when fed the same (uniform) evaluator from outside."""
    from chinesecheckersagent_b200.engine import BatchedMCTS
    roots = dev(cfg4_roots(128))
    n = roots.shape[1]
    m = BatchedMCTS(eng, num_itr=60)
    p = torch.full((n, 294), 1 / 294., dtype=torch.float64, device="cuda")
    v = torch.zeros((n,), dtype=torch.float64, device="cuda")
    a = m.search(roots, evaluator=0, pre_expand=True)
    b = m.search_with(roots, lambda leaf: (p, v), pre_expand=True)
    assert torch.equal(a["visits"], b["visits"]) and torch.equal(a["q"], b["q"]) and torch.equal(a["n_nodes"], b["n_nodes"])


def test_full_size_cfg4_properties(eng):
    """4,096 trees x 175 simulations (BASELINE configs[3]): sum N, shard invariance, spot check vs oracle."""
    from chinesecheckersagent_b200.engine import BatchedMCTS
    roots = cfg4_roots(4096)
    m = BatchedMCTS(eng)
    out = m.search(dev(roots), evaluator=0, pre_expand=False)
    visits = out["visits"].cpu().numpy()
    assert np.all(visits.sum(1) == 174)                                   # player.py:157-158 path
    assert np.all(out["n_nodes"].cpu().numpy() > 0)                        # no pool overflow
    assert np.allclose(out["pi"].cpu().numpy().sum(1), 1.0, atol=1e-12)
    sub = m.search(dev(roots[:, 1000:1100]), evaluator=0, pre_expand=False)
    assert np.array_equal(sub["visits"].cpu().numpy(), visits[1000:1100])
    v, _, _, _ = orc.mcts(roots[:, :64], 175, 3.5, 1.0, 0, 0, nthreads=8)
    assert np.array_equal(visits[:64].astype(np.uint32), v)


def test_edge_pool_overflow_is_flagged(eng):
    from chinesecheckersagent_b200.engine import BatchedMCTS
    roots = dev(cfg4_roots(8))
    out = BatchedMCTS(eng, edges_per_tree=200).search(roots)
    assert np.all(out["n_nodes"].cpu().numpy() == -1)


@pytest.mark.parametrize("evaluator,pre_expand", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_reference_tie_rule_vs_oracle(eng, evaluator, pre_expand):
    """ccx_mcts_set_tiebreak mode 1 = MCTS.py:65-72 as written (the first maximal edge plus later edges within EPSILON,
    one of them drawn uniformly) with Philox standing in for random.choice.  The C oracle builds the chosen_edges list
    literally and draws with the same counter, so N, Q and pi must agree bit for bit — in the persistent kernel and in the
    round-based pipeline — and must differ from the first-maximum rule."""
    from chinesecheckersagent_b200.engine import BatchedMCTS
    roots = cfg4_roots(192, seed=77)
    m = BatchedMCTS(eng, num_itr=120, random_ties=True, tie_seed=0xC0FFEE, tie_uid0=1000)
    out = m.search(dev(roots), evaluator=evaluator, pre_expand=bool(pre_expand))
    v, pi, q, nodes = orc.mcts(roots, 120, 3.5, 1.0, pre_expand, evaluator, nthreads=8, ties=(0xC0FFEE, 1000))
    assert np.array_equal(out["visits"].cpu().numpy().astype(np.uint32), v)
    assert np.array_equal(out["q"].cpu().numpy(), q)
    assert np.array_equal(out["pi"].cpu().numpy(), pi)
    assert np.array_equal(out["n_nodes"].cpu().numpy(), nodes)
    again = m.search(dev(roots), evaluator=evaluator, pre_expand=bool(pre_expand))       # no dependence on the handle's history
    assert torch.equal(again["visits"], out["visits"])
    first = BatchedMCTS(eng, num_itr=120).search(dev(roots), evaluator=evaluator, pre_expand=bool(pre_expand))
    assert not torch.equal(first["visits"], out["visits"])
    if evaluator == 0:
        n = roots.shape[1]
        p = torch.full((n, 294), 1 / 294., dtype=torch.float64, device="cuda")
        z = torch.zeros((n,), dtype=torch.float64, device="cuda")
        rb = m.search_with(dev(roots), lambda leaf: (p, z), pre_expand=bool(pre_expand))
        assert torch.equal(rb["visits"], out["visits"]) and torch.equal(rb["q"], out["q"])


def test_uniform_ties_spread_the_first_visits(eng):
    """With the uniform stub every edge of a fresh node ties (N_sum = 0 => U = 0): under the reference rule the first
    visit below the root is uniform over the root's edges, under the first-maximum rule it is always edge 0."""
    from chinesecheckersagent_b200.engine import BatchedMCTS
    roots = orc.start_states(4096)
    out = BatchedMCTS(eng, num_itr=2, random_ties=True, tie_seed=9).search(dev(roots), evaluator=0)
    v = out["visits"].cpu().numpy()
    assert np.all(v.sum(1) == 1)
    counts = v.sum(0)[np.nonzero(v.sum(0))[0]]
    assert len(counts) == 10                                              # the start position has 10 legal moves
    chi2 = float(((counts - 409.6) ** 2 / 409.6).sum())
    assert chi2 < 33.7, (chi2, counts)                                    # 9 degrees of freedom, p = 1e-4


@pytest.mark.parametrize("pre_expand", [0, 1])
def test_stub_search_matches_reference_on_256_roots(eng, pre_expand):
    """cfg 4 against the REFERENCE itself (not the C oracle): 256 roots searched by the unmodified MCTS.py
    (tests/golden/gen_golden_mcts256.py), both entry paths; per-edge N, root Q and pi bit for bit."""
    from chinesecheckersagent_b200.engine import BatchedMCTS
    g = dict(np.load(os.path.join(GOLDEN, "mcts_golden_256.npz")))
    assert g["roots"].shape[1] >= 256
    out = BatchedMCTS(eng).search(dev(g["roots"]), evaluator=0, pre_expand=bool(pre_expand))
    assert np.array_equal(out["visits"].cpu().numpy().astype(np.uint32), g["visits%d" % pre_expand])
    assert np.array_equal(out["q"].cpu().numpy(), g["q%d" % pre_expand])
    assert np.array_equal(out["pi"].cpu().numpy(), g["pi%d" % pre_expand])
    assert np.array_equal(out["n_nodes"].cpu().numpy(), g["nodes%d" % pre_expand])
