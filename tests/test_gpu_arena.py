"""GPU tests of the arena / evaluation path (§8f f2): Game.start bookkeeping, AiPlayer move sampling, matches."""
import os

import numpy as np
import pytest
import torch

import oracle as orc
from conftest import GOLDEN

pytestmark = pytest.mark.gpu
WEIGHTS = os.path.join(GOLDEN, "good_model_weights.npz")


@pytest.fixture(scope="module")
def eng():
    from chinesecheckersagent_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


@pytest.fixture(scope="module")
def model(eng):
    from chinesecheckersagent_b200.model import ResidualCNN
    return ResidualCNN(engine=eng).load_weights(WEIGHTS)


def test_greedy_plies_equal_play_greedy(eng):
    """ccx_game_advance in greedy mode, one ply at a time, is Game('greedy','greedy').start(): same final states and
    outcomes as k_play_greedy, which is pinned to the reference's games (tests/golden/greedy_games.npz)."""
    from chinesecheckersagent_b200.engine import BatchedEnv, _p
    n, seed = 4096, 1234
    a = BatchedEnv(n, engine=eng, seed=seed, game_id0=17)
    a.play_greedy()
    b = BatchedEnv(n, engine=eng, seed=seed, game_id0=17)
    counters = eng.zeros((4,), torch.int64)
    for _ in range(400):
        eng.call("ccx_game_advance", n, _p(b.state), None, None, seed, 17, 0.01, 16, 0, _p(counters))
    sa, sb = a.numpy_state(), b.numpy_state()
    assert np.array_equal(sa[:7], sb[:7])
    c = counters.cpu().tolist()
    status = sa[4] >> np.uint64(56)
    assert c[1] == int((status == 1).sum()) and c[2] == int((status == 2).sum()) and c[3] == int((status > 2).sum())
    assert c[1] + c[2] + c[3] == n


def test_ai_move_sampling_rule(eng):
    """tau = 0.01 -> N^100: the most visited action (player.py:151-154, MCTS.py:131-140); tau = 1 -> proportional to N."""
    from chinesecheckersagent_b200.engine import BatchedEnv, _p
    n = 20000
    env = BatchedEnv(n, engine=eng, seed=3)
    masks = env.movegen().cpu().numpy().view(np.uint64)                 # start position: ten legal moves
    legal = [k * 49 + (c >> 3) * 7 + (c & 7) for k in range(6) for c in range(56) if (int(masks[k, 0]) >> c) & 1]
    vis = np.zeros((n, 294), dtype=np.uint32)
    counts = np.array([60, 40, 30, 20, 10, 5, 4, 3, 1, 1], dtype=np.uint32)
    vis[:, legal] = counts
    visits = torch.from_numpy(vis.view(np.int32)).cuda()
    nodes = torch.ones(n, dtype=torch.int32, device="cuda")
    for tau, check in ((0.01, "argmax"), (1.0, "proportional")):
        env.reset()
        counters = eng.zeros((4,), torch.int64)
        eng.call("ccx_game_advance", n, _p(env.state), _p(visits), _p(nodes), 99, 0, tau, 16, 0, _p(counters))
        st = env.numpy_state()
        # the move played = (from, to) of the last move in meta -> policy index
        to = (st[4] >> np.uint64(8)) & np.uint64(0xFF)
        frm = st[4] & np.uint64(0xFF)
        start_cells = [48, 40, 49, 32, 41, 50]
        ids = np.array([start_cells.index(int(f)) for f in frm])
        idx = ids * 49 + (to.astype(np.int64) >> 3) * 7 + (to.astype(np.int64) & 7)
        if check == "argmax":
            assert np.all(idx == legal[0])
        else:
            freq = np.array([(idx == a).mean() for a in legal])
            assert np.abs(freq - counts / counts.sum()).max() < 0.012


def test_model_beats_greedy_and_self_match_is_sane(model):
    """SURVEY §8c behavioural anchor: good_model with the reference's MCTS settings beat the greedy player 11 of 12."""
    from chinesecheckersagent_b200.arena import GREEDY, BatchedArena, evaluate
    a = BatchedArena(model, GREEDY, 48, seed=5, num_itr=175).play()
    b = BatchedArena(GREEDY, model, 48, seed=6, game_id0=48, num_itr=175).play()
    ai = a["p1_wins"] + b["p2_wins"]
    greedy = a["p2_wins"] + b["p1_wins"]
    assert a["unfinished"] == 0 and b["unfinished"] == 0
    assert ai >= 0.7 * 96, (a, b)
    assert ai + greedy + a["stopped"] + b["stopped"] == 96
    cur, best, draws = evaluate(model, model, num_games=24)
    assert cur + best + draws == 24


def test_empty_batches_and_bad_arguments(eng):
    """every new entry point: n = 0 is a no-op returning CCX_OK, malformed calls return CCX_ERR_ARG (never crash)"""
    L, h = eng.L, eng.h
    assert L.ccx_game_advance(h, 0, None, None, None, 1, 0, 1.0, 16, 0, None) == 0
    assert L.ccx_greedy_generate(h, 0, None, 0, 1, 0, 0, 400, 43, None, None, None, 0, None, None, None, None) == 0
    assert L.ccx_cand_to_pi(h, 0, None, None) == 0
    ERR_ARG = L.ccx_game_advance(h, 4, None, None, None, 1, 0, 1.0, 16, 0, None)
    assert ERR_ARG < 0
    assert L.ccx_game_advance(h, 4, None, None, None, 1, 0, -1.0, 16, 0, None) == ERR_ARG          # tau must be positive
    assert L.ccx_greedy_generate(h, 4, None, 0, 1, 0, 0, 400, 43, None, None, None, 0, None, None, None, None) == ERR_ARG
    assert L.ccx_greedy_generate(h, 4, None, 0, 1, 0, 0, 0, 43, None, None, None, 0, None, None, None, None) == ERR_ARG
    assert L.ccx_cand_to_pi(h, 3, None, None) == ERR_ARG
    assert L.ccx_mcts_run_net(h, 0, 5, 3.5, None, 0, 0) in (0, ERR_ARG)       # ERR_ARG before any ccx_mcts_begin on this handle


def test_single_game_arena_and_one_record_generator(model):
    """smallest sizes: one arena game to the end, one generated game"""
    from chinesecheckersagent_b200.arena import GREEDY, BatchedArena
    from chinesecheckersagent_b200.data_generators import BatchedGreedyGenerator
    res = BatchedArena(GREEDY, model, 1, seed=1, num_itr=20).play()
    assert res["unfinished"] == 0 and res["p1_wins"] + res["p2_wins"] + res["stopped"] == 1
    out = BatchedGreedyGenerator(model.eng, seed=2).generate(1)
    assert out["lengths"].shape == (1,) and out["board_x"].shape[0] == int(out["lengths"][0]) > 10
