"""The Python mirror of the reference surface (Board / utils / MCTS / player / game / selfplay), used the
way selfplay.py, player.py and game.py use the reference's objects.  GPU (there is no CPU path)."""
import copy
import os

import numpy as np
import pytest

import oracle as orc
from conftest import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def first_maximum_ties():
    """the oracle comparisons below pin the deterministic tie rule; the reference-like random rule has its own tests"""
    from chinesecheckersagent_b200.MCTS import MCTS
    old, MCTS.TIE_RULE = MCTS.TIE_RULE, "first"
    yield
    MCTS.TIE_RULE = old


def test_board_start_position_known_answer():
    from chinesecheckersagent_b200.board import Board
    b = Board()
    assert b.get_valid_moves(1) == {(6, 0): [], (5, 0): [(3, 0), (5, 2)], (6, 1): [(4, 1), (6, 3)],
                                    (4, 0): [(3, 0), (4, 1)], (5, 1): [(4, 1), (5, 2)], (6, 2): [(5, 2), (6, 3)]}
    assert b.check_win() == 0 and b.player_progress(1) == 0 and b.player_forward_distance(1) == 0
    assert b.board.shape == (7, 7, 3) and b.board.dtype == np.uint8
    assert b.board[:, :, 0].tolist() == [[0, 0, 0, 0, 2, 2, 2], [0, 0, 0, 0, 0, 2, 2], [0, 0, 0, 0, 0, 0, 2], [0] * 7,
                                         [1, 0, 0, 0, 0, 0, 0], [1, 1, 0, 0, 0, 0, 0], [1, 1, 1, 0, 0, 0, 0]]


def test_board_replays_a_reference_game(env_golden):
    """First random-walk game of the fixture (80 consecutive plies recorded from the reference)."""
    from chinesecheckersagent_b200 import utils
    from chinesecheckersagent_b200.board import Board
    g = env_golden
    b = Board()
    player = 1
    for i in range(80):
        vm = b.get_valid_moves(player)
        for cid in range(6):
            want = sorted(int(x) for x in g["ref_moves"][i, cid, :g["ref_nmoves"][i, cid]])
            got = [8 * r + c for r, c in vm[b.checkers_pos[player][cid]]]
            assert got == want
        assert np.array_equal(utils.to_model_input(b, player).astype(np.uint8), g["planes"][i])
        assert np.array_equal(b.packed_state(player)[:7], g["state"][:7, i])
        f, t = int(g["chosen"][i, 0]), int(g["chosen"][i, 1])
        snap = copy.deepcopy(b)
        winner = b.place(player, (f >> 3, f & 7), (t >> 3, t & 7))
        assert winner == g["winner"][i]
        assert snap.checkers_pos != b.checkers_pos and len(snap.hist_moves) == len(b.hist_moves) - (0 if i >= 16 else 1) or i >= 16
        player = 3 - player
    assert len(b.hist_moves) == 16 and b.board[:, :, 1].any() and b.board[:, :, 2].any()


def test_index_helpers_round_trip():
    from chinesecheckersagent_b200 import board_utils, utils
    for i in range(294):
        cid, pos = utils.decode_checker_index(i)
        assert utils.encode_checker_index(cid, pos) == i                      # model.py:182-191 intent
    for r in range(7):
        for c in range(7):
            assert board_utils.human_coord_to_np_index(board_utils.np_index_to_human_coord((r, c))) == (r, c)


class StubModel:
    version = 0

    def predict(self, x):
        assert x.shape == (7, 7, 7)
        return np.full(294, 1 / 294.), 0.0


def test_mcts_object_api_matches_oracle():
    from chinesecheckersagent_b200.board import Board
    from chinesecheckersagent_b200.MCTS import MCTS, Node
    root = Node(Board(), 1)
    np.random.seed(0)
    pi, edge = MCTS(root, StubModel(), num_itr=60).search()              # player.py:157-158 path
    v, opi, oq, _ = orc.mcts(orc.start_states(1), 60, 3.5, 1.0, 0, 0)
    assert np.array_equal(pi, opi[0]) and pi.dtype == np.float64
    assert sum(e.stats['N'] for e in root.edges) == 59
    assert edge.outNode.state.hist_moves[-1] == (edge.fromPos, edge.toPos)
    for e in root.edges:
        cid = root.state.checkers_id[1][e.fromPos]
        idx = cid * 49 + e.toPos[0] * 7 + e.toPos[1]
        assert e.stats['N'] == v[0, idx] and e.stats['Q'] == oq[0, idx]


def test_make_move_flow_with_noise_matches_oracle():
    """selfplay.py:114-127: expand root, mix Dirichlet noise into stats['P'], search."""
    from chinesecheckersagent_b200.board import Board
    from chinesecheckersagent_b200.MCTS import MCTS, Node
    root = Node(Board(), 1)
    tree = MCTS(root, StubModel(), num_itr=50)
    tree.expandAndBackUp(tree.root, breadcrumbs=[])
    assert len(root.edges) == 10
    noise = np.random.default_rng(3).dirichlet(np.ones(10) * 0.03)
    for i in range(10):
        root.edges[i].stats['P'] *= (1. - 0.25)
        root.edges[i].stats['P'] += 0.25 * noise[i]
    pi, _ = tree.search()
    nz = np.zeros((1, 16)); nz[0, :10] = noise
    v, opi, _, _ = orc.mcts(orc.start_states(1), 50, 3.5, 1.0, 1, 0, noise=nz)
    assert np.array_equal(pi, opi[0]) and sum(e.stats['N'] for e in root.edges) == 50


def test_greedy_player_and_game(env_golden):
    from chinesecheckersagent_b200.board import Board
    from chinesecheckersagent_b200.game import Game
    from chinesecheckersagent_b200.player import GreedyPlayer
    cands = GreedyPlayer(1).decide_move(Board(), training=True)
    g = env_golden
    want = set()
    for k in range(g["n_greedy"][0]):
        f, t = int(g["greedy"][0, k, 0]), int(g["greedy"][0, k, 1])
        want.add(((f >> 3) - (f & 7) + 7, min(f >> 3, f & 7) + 1, (t >> 3) - (t & 7) + 7, min(t >> 3, t & 7) + 1))
    assert {(s[0], s[1], e[0], e[1]) for s, e in cands} == want
    results = [Game(p1_type='greedy', p2_type='greedy', verbose=False).start() for _ in range(5)]
    assert all(r in (1, 2, None) for r in results) and any(r in (1, 2) for r in results)


def test_selfplay_function_contract():
    from chinesecheckersagent_b200 import utils
    from chinesecheckersagent_b200.engine import Engine
    from chinesecheckersagent_b200.board import default_engine
    from chinesecheckersagent_b200.model import ResidualCNN
    from chinesecheckersagent_b200.selfplay import selfplay
    model = ResidualCNN(engine=default_engine()).load_weights(os.path.join(GOLDEN, "good_model_weights.npz"))
    np.random.seed(1)
    games = []
    for _ in range(3):
        hist, reward = selfplay(model, num_itr=12)
        if hist is not None:
            games.append((hist, reward))
            assert reward in (1, -1) and all(pi.shape == (294,) and abs(pi.sum() - 1) < 1e-9 for _, pi in hist)
    if games:
        bx, py, vy = utils.convert_to_train_data(games)                      # train.py:269
        assert len(bx) == len(py) == len(vy) and bx[0].shape == (7, 7, 7)
        assert vy[0] == games[0][1] and (len(vy) < 2 or vy[1] == -vy[0])


def test_stochastic_greedy_branch_samples_legal_forward_moves():
    """GreedyPlayer(stochastic=True) (player.py:77-97): every pick is legal; forward moves are drawn in proportion to the rows
    they advance, so the empirical mean advance from the start position sits at sum(d^2)/sum(d) over the forward moves."""
    import random

    from chinesecheckersagent_b200 import board_utils
    from chinesecheckersagent_b200.board import Board
    from chinesecheckersagent_b200.player import GreedyPlayer
    random.seed(3); np.random.seed(3)
    b = Board()
    valid = b.get_valid_moves(1)
    human = board_utils.convert_np_to_human_moves(valid)
    dists = [s[0] - e[0] for s in human for e in human[s] if s[0] - e[0] > 0]
    p = GreedyPlayer(1, stochastic=True)
    adv = []
    for _ in range(400):
        s, e = p.decide_move(b)
        assert e in valid[s]
        adv.append(board_utils.np_index_to_human_coord(s)[0] - board_utils.np_index_to_human_coord(e)[0])
    assert min(adv) > 0
    want = sum(d * d for d in dists) / sum(dists)
    assert abs(np.mean(adv) - want) < 0.15


def test_native_search_path_equals_the_per_simulation_round_trips():
    """MCTS(...).search() with the package's own net runs all simulations inside libccx (ccx_mcts_run_net, graph replay);
    the visit counts must be the ones the per-simulation select / predict / expand_backup round trips produce."""
    from chinesecheckersagent_b200.board import Board, default_engine
    from chinesecheckersagent_b200.MCTS import MCTS, Node
    from chinesecheckersagent_b200.model import ResidualCNN
    model = ResidualCNN(engine=default_engine()).load_weights(os.path.join(GOLDEN, "good_model_weights.npz"))

    class Foreign:                                   # same arithmetic, but opaque to the fast path: forces the round trips
        version = 0

        def predict(self, x):
            return model.predict(x)
    for flow in ("decide_move", "make_move"):
        out = []
        for m in (model, Foreign(), model, model):   # the third / fourth native search replay the cached graph
            np.random.seed(4)
            tree = MCTS(Node(Board(), 1), m, num_itr=40)
            assert tree._native == (m is model)
            if flow == "make_move":
                tree.expandAndBackUp(tree.root, breadcrumbs=[])
            pi, edge = tree.search()
            out.append((pi.copy(), [e.stats['N'] for e in tree.root.edges], [e.stats['W'] for e in tree.root.edges]))
        for o in out[1:]:
            assert np.array_equal(out[0][0], o[0]) and out[0][1] == o[1]
            assert np.allclose(out[0][2], o[2], rtol=0, atol=1e-12)
        assert sum(out[0][1]) == (39 if flow == "decide_move" else 40)


def test_random_tie_rule_is_seeded_by_python_random_and_spreads_first_visits():
    """MCTS.TIE_RULE = "random" (the default): ties within EPSILON are drawn uniformly (MCTS.py:65-72).  With the uniform
    stub every edge of a fresh node ties, so the first visits must not all go to edge 0, and `random.seed` must make the
    search reproducible like it does for the reference's random.choice."""
    import random

    from chinesecheckersagent_b200.board import Board
    from chinesecheckersagent_b200.MCTS import MCTS, Node
    MCTS.TIE_RULE = "random"
    runs = []
    for seed in (11, 11, 12):
        random.seed(seed); np.random.seed(0)
        root = Node(Board(), 1)
        MCTS(root, StubModel(), num_itr=8).search()
        runs.append([e.stats['N'] for e in root.edges])
    assert runs[0] == runs[1] and sum(runs[0]) == 7
    firsts = set()
    for seed in range(12):
        random.seed(100 + seed); np.random.seed(0)
        root = Node(Board(), 1)
        MCTS(root, StubModel(), num_itr=2).search()          # simulation 1 expands the root, simulation 2 visits one edge
        firsts.add([e.stats['N'] for e in root.edges].index(1))
    assert len(firsts) >= 4, firsts
