"""GPU tests of the batched self-play pipeline (selfplay.py semantics) through the C-ABI."""
import os

import numpy as np
import pytest
import torch

import oracle as orc
import selfplay_ref
from conftest import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from chinesecheckersagent_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


def run_selfplay(eng, n=256, iters=140, sims=12, seed=77, **kw):
    from chinesecheckersagent_b200.selfplay import BatchedSelfPlay, UniformEvaluator
    sp = BatchedSelfPlay(eng, UniformEvaluator(eng, n), n_slots=n, seed=seed, num_itr=sims, max_iters=iters, log_moves=True, **kw)
    sp.run(iters=iters)
    return sp


def games_from_log(sp):
    """Split every slot's move log into game instances."""
    n = sp.n
    log = sp.move_log.cpu().numpy().view(np.uint32).reshape(sp.max_iters, n)
    games = []
    for g in range(n):
        cur = []
        for it in range(sp.iter):
            w = int(log[it, g])
            if not (w >> 24) & 1:
                continue
            cur.append((w & 0xFF, (w >> 8) & 0xFF, (w >> 16) & 0xFF, (w >> 25) & 1, it))
            if (w >> 16) & 0xFF:
                games.append((g, cur)); cur = []
    return games


def test_bookkeeping_matches_reference_rules(eng):
    """Every game the GPU played, replayed through the restatement of selfplay.py:29-80: same opening
    length, same status after every ply, same tau switch, same records kept/dropped."""
    sp = run_selfplay(eng)
    games = games_from_log(sp)
    assert len(games) > 50
    flags = sp.rec_flag.cpu().numpy().reshape(sp.max_iters, sp.n)
    statuses = set()
    for slot, moves in games:
        out, _ = selfplay_ref.replay([(m[0], m[1]) for m in moves])
        assert len(out) == len(moves)
        for o, m in zip(out, moves):
            assert o["status"] == m[2] and o["mcts"] == bool(m[3])
            f = int(flags[m[4], slot])
            if o["mcts"]:
                assert bool(f & 0x10) == o["tau_det"]
                final = moves[-1][2]
                assert (f & 0xF) in ((2, 3) if final in (1, 2) else (4,))
            else:
                assert f == 0
        statuses.add(moves[-1][2])
    assert {3, 4} & statuses                      # near-random play: discards happen
    st = sp.stats()
    assert st["p1_wins"] + st["p2_wins"] + st["discarded_repetition"] + st["discarded_no_progress"] == len(games)


def test_trajectory_format_and_contents(eng):
    from chinesecheckersagent_b200.model import ResidualCNN
    from chinesecheckersagent_b200.selfplay import BatchedSelfPlay
    model = ResidualCNN(engine=eng).load_weights(os.path.join(GOLDEN, "good_model_weights.npz"))
    sp = BatchedSelfPlay(eng, model.evaluate_states, n_slots=256, num_itr=24, max_iters=130, seed=5)
    sp.run(iters=130)
    traj = sp.collect()
    bx, pi, vy = traj["board_x"].cpu().numpy(), traj["pi_y"].cpu().numpy(), traj["v_y"].cpu().numpy()
    state = traj["state"].cpu().numpy().view(np.uint64)
    m = bx.shape[0]
    assert m == sp.stats()["records"]
    st = sp.stats()
    assert st["p1_wins"] + st["p2_wins"] > 0 and m > 0, st     # the net finishes games
    assert bx.shape == (m, 7, 7, 7) and pi.shape == (m, 294) and set(np.unique(vy)) <= {-1, 1}
    full = np.zeros((8, m), dtype=np.uint64); full[:5] = state
    assert np.array_equal(bx, orc.encode(full))                                  # utils.to_model_input of the root state
    assert np.allclose(pi.sum(1), 1.0, atol=1e-5)
    # the reward alternates along a game and the first recorded ply is ply 6 with player 1 to move
    plies = ((state[4] >> np.uint64(32)) & np.uint64(0xFFFF)).astype(int)
    to_move = ((state[4] >> np.uint64(48)) & np.uint64(1)).astype(int)
    assert plies.min() == 6 and np.all(to_move == (plies & 1))
    assert np.all(to_move[plies == 6] == 0)
    masks = orc.movegen(full)                                                    # pi is supported on legal moves only
    for i in range(0, m, max(1, m // 200)):
        for a in np.nonzero(pi[i])[0]:
            cid, off = divmod(int(a), 49)
            assert (int(masks[cid, i]) >> ((off // 7) * 8 + off % 7)) & 1


def test_same_seed_same_games(eng):
    a = run_selfplay(eng, n=128, iters=40, sims=6, seed=9)
    b = run_selfplay(eng, n=128, iters=40, sims=6, seed=9)
    assert torch.equal(a.move_log, b.move_log) and torch.equal(a.rec_visits, b.rec_visits)
    c = run_selfplay(eng, n=128, iters=40, sims=6, seed=10)
    assert not torch.equal(a.move_log, c.move_log)


def test_gamma_noise_is_dirichlet_like(eng):
    import ctypes
    n, stride = 4096, 32
    out = torch.empty((n, stride), dtype=torch.float64, device="cuda")
    eng.call("ccx_gamma_noise", n, stride, 0.03, 123, 0, 0, ctypes.c_void_p(out.data_ptr()))
    g = out.cpu().numpy()
    assert np.all(g >= 0) and np.all(np.isfinite(g))
    # Gamma(alpha, 1): mean alpha, variance alpha
    assert abs(g.mean() - 0.03) < 0.003 and abs(g.var() - 0.03) < 0.006
    d = g / g.sum(1, keepdims=True)
    assert abs(d.mean() - 1 / stride) < 1e-6
    # Dirichlet(0.03 * 1_32) against numpy's sampler (20k draws): median of the largest coordinate 0.630,
    # mean number of coordinates above 0.01 is 4.23
    assert 0.60 < np.median(d.max(1)) < 0.66
    assert 3.9 < (d > 0.01).sum(1).mean() < 4.6


def test_selfplay_with_real_net_smoke(eng):
    from chinesecheckersagent_b200.model import ResidualCNN
    from chinesecheckersagent_b200.selfplay import BatchedSelfPlay
    model = ResidualCNN(engine=eng).load_weights(os.path.join(GOLDEN, "good_model_weights.npz"))
    sp = BatchedSelfPlay(eng, model.evaluate_states, n_slots=64, num_itr=16, max_iters=12)
    st = sp.run(iters=10)
    assert st["plies"] == 64 * 10
    v = sp.visits.cpu().numpy()
    assert np.all(v.sum(1) == 16)                      # root pre-expanded: sum N = num_itr (selfplay.py:117,127)


def test_full_selfplay_with_real_net_at_size(eng):
    """cfg 5 at working size: 1,024 slots play whole games with the reference's settings (175 simulations, Dirichlet noise,
    tau switch) through the fused net-MCTS rounds; the trajectory buffer must be consistent with the outcome counters."""
    from chinesecheckersagent_b200.model import ResidualCNN
    from chinesecheckersagent_b200.selfplay import BatchedSelfPlay
    model = ResidualCNN(engine=eng).load_weights(os.path.join(GOLDEN, "good_model_weights.npz"))
    sp = BatchedSelfPlay(eng, model.evaluate_states, n_slots=1024, max_iters=80, seed=2026)
    assert sp.fused
    st = sp.run(iters=80)
    assert st["plies"] == 1024 * 80                      # every slot plays every iteration (finished slots restart at once)
    finished = st["p1_wins"] + st["p2_wins"]
    assert finished > 300 and st["discarded_overflow"] == 0
    assert finished == st["games"]
    assert 0.25 < st["p1_wins"] / finished < 0.75
    traj = sp.collect()
    m = traj["board_x"].shape[0]
    assert m == st["records"] and traj["pi_y"].shape == (m, 294) and traj["v_y"].shape == (m,)
    pi = traj["pi_y"]
    assert torch.allclose(pi.sum(1), torch.ones(m, device=pi.device), atol=1e-4)
    v = traj["v_y"].cpu().numpy()
    assert set(np.unique(v)) <= {-1, 1} and 0.3 < (v == 1).mean() < 0.7
    # records are positions after the six random opening plies: planes 0/1 hold six labelled checkers each
    bx = traj["board_x"].cpu().numpy()
    assert np.all((bx[..., 0] > 0).sum((1, 2)) == 6) and np.all((bx[..., 1] > 0).sum((1, 2)) == 6)


def test_selfplay_records_do_not_depend_on_the_trunk_kernel(eng):
    """the accurate trunk's kernel choice (one tile per CTA, or 1-3 contexts sharing the weight slots; automatic by default) is
    invisible above the net: whole self-play runs (fused rounds, graph replay, restarts) deliver the same records bit for bit"""
    from chinesecheckersagent_b200.model import ResidualCNN
    from chinesecheckersagent_b200.selfplay import BatchedSelfPlay
    model = ResidualCNN(engine=eng).load_weights(os.path.join(GOLDEN, "good_model_weights.npz"))
    model.set_kernel("tc_acc")
    out = []
    try:
        for ctx in (0, 3, -1):
            eng.call("ccx_net_set_acc_contexts", ctx)
            sp = BatchedSelfPlay(eng, model.evaluate_states, n_slots=700, num_itr=10, max_iters=140, seed=99)    # 175 tiles: contexts without a tile
            st = sp.run(iters=140)
            traj = sp.collect()
            out.append((st["plies"], st["records"], traj["board_x"].clone(), traj["pi_y"].clone(), traj["v_y"].clone(), sp.visits.clone()))
    finally:
        eng.call("ccx_net_set_acc_contexts", -1)
    for o in out[1:]:
        assert o[0] == out[0][0] and o[1] == out[0][1]
        assert torch.equal(o[2], out[0][2]) and torch.equal(o[3], out[0][3]) and torch.equal(o[4], out[0][4]) and torch.equal(o[5], out[0][5])
    assert out[0][1] > 1000


def test_collect_consumes_its_records(eng):
    from chinesecheckersagent_b200.model import ResidualCNN
    from chinesecheckersagent_b200.selfplay import BatchedSelfPlay
    model = ResidualCNN(engine=eng).load_weights(os.path.join(GOLDEN, "good_model_weights.npz"))
    sp = BatchedSelfPlay(eng, model.evaluate_states, n_slots=128, num_itr=16, max_iters=110, seed=3)
    sp.run(iters=110)
    a = sp.collect()
    assert a["v_y"].shape[0] == sp.stats()["records"] > 0
    b = sp.collect()
    assert b["v_y"].shape[0] == 0 and b["board_x"].shape == (0, 7, 7, 7)


def test_ring_buffer_plays_past_max_iters(eng):
    """ring=True: the record buffers hold 48 iterations but the loop runs 200; finished records are moved out in time, games
    longer than max_iters // 2 iterations are discarded as overflow, and every kept record is delivered exactly once."""
    from chinesecheckersagent_b200.model import ResidualCNN
    from chinesecheckersagent_b200.selfplay import BatchedSelfPlay, UniformEvaluator
    model = ResidualCNN(engine=eng).load_weights(os.path.join(GOLDEN, "good_model_weights.npz"))
    n = 192
    sp = BatchedSelfPlay(eng, model.evaluate_states, n_slots=n, seed=21, num_itr=12, max_iters=160, ring=True)
    st = sp.run(iters=400)
    assert st["iterations"] == 400 and st["plies"] > 0
    traj = sp.collect()
    assert traj["v_y"].shape[0] == st["records"] > 0
    assert sp.collect()["v_y"].shape[0] == 0
    ended = st["p1_wins"] + st["p2_wins"] + st["discarded_repetition"] + st["discarded_no_progress"] + st["discarded_overflow"]
    assert ended > n                                          # slots restarted
    state = traj["state"].cpu().numpy().view(np.uint64)
    full = np.zeros((8, state.shape[1]), dtype=np.uint64); full[:5] = state
    assert np.array_equal(traj["board_x"].cpu().numpy(), orc.encode(full))
    # a linear buffer of the same size refuses to go past its end
    lin = BatchedSelfPlay(eng, UniformEvaluator(eng, n), n_slots=n, seed=21, num_itr=8, max_iters=6)
    lin.run(iters=6)
    with pytest.raises(RuntimeError):
        lin.step()


def test_play_games_starts_exactly_the_requested_games(eng):
    """train.py:58-64: exactly num_self_play games are started and every one is played out (kept or discarded); nothing is
    left in flight and no slot plays extra games."""
    from chinesecheckersagent_b200.model import ResidualCNN
    from chinesecheckersagent_b200.selfplay import BatchedSelfPlay
    model = ResidualCNN(engine=eng).load_weights(os.path.join(GOLDEN, "good_model_weights.npz"))
    for n, games in ((64, 150), (64, 40), (32, 32)):
        sp = BatchedSelfPlay(eng, model.evaluate_states, n_slots=n, seed=5, num_itr=12, max_iters=400)
        st = sp.play_games(games)
        ended = st["p1_wins"] + st["p2_wins"] + st["discarded_repetition"] + st["discarded_no_progress"] + st["discarded_overflow"]
        assert st["unfinished"] == 0 and st["games_started"] == games and ended == games, st
        assert st["games"] == st["p1_wins"] + st["p2_wins"] > 0
        assert int(sp.serial.sum().item()) == games - min(n, games)            # restarts = games beyond the first one per slot
        assert sp.collect()["v_y"].shape[0] == st["records"]


def test_two_net_selfplay_with_twin_nets_equals_single_net(eng):
    """selfplay(model1, model2) (selfplay.py:11-29,58): model1 searches the even plies, model2 the odd ones.  With the same
    weights loaded into two engines the split search must reproduce the single-net games bit for bit."""
    from chinesecheckersagent_b200.engine import Engine
    from chinesecheckersagent_b200.model import ResidualCNN
    from chinesecheckersagent_b200.selfplay import BatchedSelfPlay
    w = os.path.join(GOLDEN, "good_model_weights.npz")
    m1 = ResidualCNN(engine=eng).load_weights(w)
    eng2 = Engine(0)
    m2 = ResidualCNN(engine=eng2).load_weights(w)
    kw = dict(n_slots=96, num_itr=16, max_iters=40, seed=13, log_moves=True)
    a = BatchedSelfPlay(eng, m1.evaluate_states, **kw)
    a.run(iters=40)
    b = BatchedSelfPlay(eng, m1.evaluate_states, opponent=m2, **kw)
    b.run(iters=40)
    assert torch.equal(a.move_log, b.move_log) and torch.equal(a.rec_visits, b.rec_visits) and torch.equal(a.rec_flag, b.rec_flag)
    assert a.stats() == b.stats()
    with pytest.raises(ValueError):
        BatchedSelfPlay(eng, m1.evaluate_states, opponent=m1, **kw)
    eng2.close()


def test_compaction_during_the_drain_changes_nothing(eng):
    """play_games() drops finished slots from the batch while it drains (the net stops evaluating dummy positions for them).
    The surviving games must continue bit for bit — same outcomes, same records — because every Philox stream (opening moves,
    Dirichlet noise, tie draws, move sampling) is keyed by the slot's identity, not by its position in the batch."""
    from chinesecheckersagent_b200.model import ResidualCNN
    from chinesecheckersagent_b200.selfplay import BatchedSelfPlay
    model = ResidualCNN(engine=eng).load_weights(os.path.join(GOLDEN, "good_model_weights.npz"))

    def run(compact):
        sp = BatchedSelfPlay(eng, model.evaluate_states, n_slots=320, seed=9, num_itr=10, max_iters=400)
        st = sp.play_games(400, compact=compact, compact_min=24)
        t = sp.collect()
        key = np.vstack([t["state"].cpu().numpy(), t["v_y"].cpu().numpy()[None].astype(np.int64),
                         (t["pi_y"].double() * torch.arange(1, 295, device=t["pi_y"].device)).sum(1).mul(1e6).round().long().cpu().numpy()[None]])
        order = np.lexsort(key)
        return st, t["board_x"].cpu().numpy()[order], t["pi_y"].cpu().numpy()[order], t["v_y"].cpu().numpy()[order], sp.n
    a, b = run(False), run(True)
    assert b[0]["compactions"] >= 1 and a[0]["compactions"] == 0 and b[4] < a[4] == 320
    for k in ("plies", "p1_wins", "p2_wins", "discarded_repetition", "discarded_no_progress", "discarded_overflow", "records", "games",
              "iterations", "games_started", "unfinished"):
        assert a[0][k] == b[0][k], k
    assert a[1].shape == b[1].shape and a[1].shape[0] == a[0]["records"] > 0
    assert np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3])
