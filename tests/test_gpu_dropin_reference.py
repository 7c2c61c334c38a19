"""The literal drop-in: the UNMODIFIED reference callers (selfplay.py, game.py, player.py, MCTS.py — the staged copy under
oracle/_ref/reference.zip, see oracle/refrun.py) run on top of this package with `board`, `board_utils`, `utils`, `model`
(and, in the first scenario, `MCTS`) replaced by the facade modules through sys.modules — what a maintainer gets by
putting the package in front of the reference's flat modules.  GPU only (the facade has no CPU path)."""
import contextlib
import importlib
import io
import os
import random
import sys

import numpy as np
import pytest

import oracle as orc
import refrun
from conftest import GOLDEN

pytestmark = pytest.mark.gpu

REF_MODULES = ("board", "board_utils", "utils", "model", "MCTS", "player", "game", "selfplay", "config", "loss", "data_generators")


@contextlib.contextmanager
def reference_over_facade(facade_mcts):
    """sys.modules view in which the reference's callers import the facade for everything on the hot path"""
    ref_dir = refrun.reference_dir()
    if ref_dir is None:
        pytest.skip("no staged reference (oracle/_ref/reference.zip) (run __graft_entry__.build() where /root/reference exists)")
    import chinesecheckersagent_b200.board as f_board
    import chinesecheckersagent_b200.board_utils as f_board_utils
    import chinesecheckersagent_b200.MCTS as f_mcts
    import chinesecheckersagent_b200.model as f_model
    import chinesecheckersagent_b200.utils as f_utils
    saved = {k: sys.modules.get(k) for k in REF_MODULES}
    saved_path = list(sys.path)
    try:
        for k in REF_MODULES:
            sys.modules.pop(k, None)
        sys.modules.update(board=f_board, board_utils=f_board_utils, utils=f_utils, model=f_model)
        if facade_mcts:
            sys.modules["MCTS"] = f_mcts
        sys.path.insert(0, ref_dir)                    # config, player, game, selfplay (and MCTS in scenario B) come from the reference
        sys.dont_write_bytecode = True
        mods = {name: importlib.import_module(name) for name in ("MCTS", "player", "game", "selfplay")}
        for name in ("player", "game", "selfplay") + (() if facade_mcts else ("MCTS",)):
            assert os.path.abspath(mods[name].__file__).startswith(os.path.abspath(ref_dir)), name      # the reference's file, not ours
        assert mods["game"].Board is f_board.Board and mods["selfplay"].Board is f_board.Board
        assert (mods["player"].MCTS is f_mcts.MCTS) == facade_mcts
        yield mods
    finally:
        sys.path[:] = saved_path
        for k in REF_MODULES:
            sys.modules.pop(k, None)
            if saved[k] is not None:
                sys.modules[k] = saved[k]


@pytest.fixture(scope="module")
def model():
    from chinesecheckersagent_b200.board import default_engine
    from chinesecheckersagent_b200.model import ResidualCNN
    return ResidualCNN(engine=default_engine()).load_weights(os.path.join(GOLDEN, "good_model_weights.npz"))


def test_reference_game_start_plays_greedy_games_on_the_facade_board():
    """game.py:58-100 + player.py:67-129 unmodified; Board from the package."""
    with reference_over_facade(True) as ref, contextlib.redirect_stdout(io.StringIO()):
        random.seed(7)
        results, plies = [], []
        for _ in range(5):
            g = ref["game"].Game(p1_type='greedy', p2_type='greedy', verbose=False)
            results.append(g.start())
            plies.append(g.board._plies)
            cw = g.board.check_win()
            assert (results[-1] in (1, 2) and cw == results[-1]) or (results[-1] is None and cw == 0)
    assert all(r in (1, 2, None) for r in results) and any(r in (1, 2) for r in results)
    assert 20 < np.mean(plies) < 80                      # the reference averages 43 plies per greedy game (config.py:77)


def test_reference_selfplay_function_runs_on_the_facade(model):
    """selfplay.py:11-133 unmodified — make_random_move, make_move with Dirichlet noise on edge.stats['P'], the repetition /
    progress rules — over the package's Board, MCTS (device trees, net inside libccx) and ResidualCNN."""
    from chinesecheckersagent_b200 import utils
    with reference_over_facade(True) as ref, contextlib.redirect_stdout(io.StringIO()):
        random.seed(3); np.random.seed(3)
        games = []
        for _ in range(4):
            hist, reward = ref["selfplay"].selfplay(model)
            if hist is not None:
                games.append((hist, reward))
                break
    assert games, "four self-play games in a row were discarded"
    hist, reward = games[0]
    assert reward in (1, -1) and len(hist) >= 4
    for board, pi in hist:
        assert pi.shape == (294,) and pi.dtype == np.float64 and abs(pi.sum() - 1) < 1e-9
        legal = board.get_valid_moves(1 if (board._plies % 2 == 0) else 2)
        mover = 1 if (board._plies % 2 == 0) else 2
        for a in np.nonzero(pi)[0]:
            cid, to = utils.decode_checker_index(int(a))
            assert to in legal[board.checkers_pos[mover][cid]]
    assert hist[0][0]._plies == 6                        # INITIAL_RANDOM_MOVES random plies come first (selfplay.py:32-33)
    bx, py, vy = utils.convert_to_train_data(games)      # train.py:269
    assert len(bx) == len(hist) and vy[0] == reward and (len(vy) < 2 or vy[1] == -reward)


def test_reference_ai_player_decides_a_move_on_the_facade(model):
    """player.py:133-166 unmodified: AiPlayer.decide_move -> MCTS(node, model, tree_tau).search() on an unexpanded root."""
    from chinesecheckersagent_b200.board import Board
    with reference_over_facade(True) as ref, contextlib.redirect_stdout(io.StringIO()):
        np.random.seed(5); random.seed(5)
        b = Board()
        ai = ref["player"].AiPlayer(player_num=1, model=model, tree_tau=0.01)
        frm, to = ai.decide_move(b, verbose=False, total_moves=0)
        assert to in b.get_valid_moves(1)[frm]
        # and a full reference Game between the AiPlayer and the reference's GreedyPlayer finishes
        g = ref["game"].Game(p1_type='ai', p2_type='greedy', verbose=False, model1=model)
        winner = g.start()
        assert winner in (1, 2, None)


class StubModel:
    version = 0

    def predict(self, x):
        assert x.shape == (7, 7, 7)
        return np.full(294, 1 / 294.), 0.0


def test_reference_mcts_py_over_the_facade_board_equals_the_device_tree():
    """Scenario B: the reference's OWN MCTS.py (Node / Edge objects, deepcopy + place per child, Python PUCT loop) running on
    the package's Board / utils.  Its visit counts, with random.choice pinned to the first candidate, must equal the CUDA
    tree's (first-maximum rule) and the C oracle's — three implementations of MCTS.py:49-137 over two implementations of
    board.py, one answer."""
    from chinesecheckersagent_b200.board import Board
    from chinesecheckersagent_b200.MCTS import MCTS as DeviceMCTS, Node as DeviceNode

    class First:
        @staticmethod
        def choice(seq):
            return seq[0]
    with reference_over_facade(False) as ref:
        M = ref["MCTS"]
        M.random = First
        np.random.seed(1)
        root = M.Node(Board(), 1)
        pi_ref, _ = M.MCTS(root, StubModel(), num_itr=40).search()
        n_ref = {(e.fromPos, e.toPos): e.stats['N'] for e in root.edges}
    old, DeviceMCTS.TIE_RULE = DeviceMCTS.TIE_RULE, "first"
    try:
        np.random.seed(1)
        droot = DeviceNode(Board(), 1)
        pi_dev, _ = DeviceMCTS(droot, StubModel(), num_itr=40).search()
    finally:
        DeviceMCTS.TIE_RULE = old
    n_dev = {(e.fromPos, e.toPos): e.stats['N'] for e in droot.edges}
    assert n_ref == n_dev and sum(n_ref.values()) == 39
    assert np.array_equal(pi_ref, pi_dev)
    _, opi, _, _ = orc.mcts(orc.start_states(1), 40, 3.5, 1.0, 0, 0)
    assert np.array_equal(pi_ref, opi[0])
