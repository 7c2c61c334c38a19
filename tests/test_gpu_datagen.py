"""GPU tests of the greedy supervised-data generator (§8f f3; data_generators.py:14-80) through the C-ABI."""
import os

import numpy as np
import pytest
import torch

import oracle as orc
from conftest import GOLDEN

pytestmark = pytest.mark.gpu
SEED = 0x5EED2026


@pytest.fixture(scope="module")
def eng():
    from chinesecheckersagent_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


def test_reference_records_through_the_kernels(eng):
    """positions recorded by the real GreedyDataGenerator -> ccx_greedy_candidates + ccx_cand_to_pi + ccx_encode
    give the reference's pi and convert_to_train_data's board_x bit for bit"""
    from chinesecheckersagent_b200.engine import BatchedEnv, _p
    from chinesecheckersagent_b200.config import DTYPE_U8
    g = dict(np.load(os.path.join(GOLDEN, "datagen_golden.npz")))
    st = g["state"]
    m = st.shape[1]
    env = BatchedEnv(m, engine=eng, state=st)
    cand = env.greedy_candidates()
    pi = eng.empty((m, 294), torch.float32)
    eng.call("ccx_cand_to_pi", m, _p(cand), _p(pi))
    assert np.array_equal(pi.cpu().numpy(), g["pi"].astype(np.float32))
    xy = g["has_xy"].astype(bool)
    assert np.array_equal(env.encode(DTYPE_U8).cpu().numpy()[xy], g["board_x"][xy])


def replay(kind, n, gid0, st0=None, stuck=400):
    """CPU restatement of generate_play with the engine's RNG rule, on the pinned oracle primitives."""
    st = st0 if st0 is not None else orc.start_states(n)
    recs = []
    for i in range(n):
        s = st[:, i:i + 1].copy()
        gid = gid0 + i
        ply = 0
        if kind == "random_start":
            s, _, _ = orc.step_random(s, SEED, 0, 6, game_id0=gid)
            ply = 6
        hist, winner = [], 0
        while True:
            cand = orc.greedy_candidates(s)
            moves = [(k, c) for k in range(6) for c in range(56) if (int(cand[k, 0]) >> c) & 1]
            if not moves or len(hist) >= stuck:
                break
            hist.append((s[:5, 0].copy(), cand[:, 0].copy()))
            r = orc.philox(SEED & 0xFFFFFFFF, SEED >> 32, ply, 1, gid & 0xFFFFFFFF, gid >> 32)
            k, c = moves[(int(r[0]) * len(moves)) >> 32]
            p2 = (int(s[4, 0]) >> 48) & 1
            frm = (int(s[3 if p2 else 2, 0]) >> (8 * k)) & 0xFF
            s, w = orc.apply(s, np.array([frm], np.uint8), np.array([c], np.uint8))
            ply += 1
            if w[0]:
                winner = int(w[0]); break
        if winner == 0:
            hist = hist[:43]
        elif kind == "randomised":
            hist = hist[3:]
        recs.append((hist, winner))
    return recs


@pytest.mark.parametrize("kind", ["normal", "random_start", "randomised"])
def test_generator_equals_cpu_restatement(eng, kind):
    from chinesecheckersagent_b200.data_generators import BatchedGreedyGenerator
    from chinesecheckersagent_b200.engine import BatchedEnv
    n = 96
    gen = BatchedGreedyGenerator(eng, seed=SEED)
    out = gen.generate(n, randomised=kind == "randomised", random_start=kind == "random_start")
    # randomised starts come from ccx_reset (validated in test_gpu_env.py); the replay starts from the same placements
    st0 = BatchedEnv(n, engine=eng, seed=SEED, game_id0=0, randomised=True).numpy_state() if kind == "randomised" else None
    want = replay(kind, n, 0, st0)
    lengths = out["lengths"].cpu().numpy()
    assert list(lengths) == [len(h) for h, _ in want]
    assert list(out["winners"].cpu().numpy()) == [w for _, w in want]
    st = out["state"].cpu().numpy().view(np.uint64)
    cand = out["cand"].cpu().numpy().view(np.uint64)
    v = out["v_y"].cpu().numpy()
    r = 0
    for hist, winner in want:
        for s, c in hist:
            assert np.array_equal(st[:, r], s) and np.array_equal(cand[:, r], c)
            mover = ((int(s[4]) >> 48) & 1) + 1
            assert v[r] == (0 if winner == 0 else (1 if mover == winner else -1))
            r += 1
    assert r == st.shape[1]
    assert np.array_equal(out["board_x"].cpu().numpy(), orc.encode(np.vstack([st, np.zeros((3, r), np.uint64)])))


def test_generator_properties_at_size(eng):
    """1e5 games of the train_on_greedy mix: every pi row sums to 1 on legal greedy candidates, lengths and outcome
    statistics sit where the reference's do (~43 plies per game, P1 ~52 %)."""
    from chinesecheckersagent_b200.data_generators import BatchedGreedyGenerator
    gen = BatchedGreedyGenerator(eng, seed=SEED)
    out = gen.generate(100000)
    pi = out["pi_y"]
    assert torch.allclose(pi.sum(1), torch.ones(pi.shape[0], device=pi.device), atol=1e-5)
    lengths = out["lengths"].float()
    assert 40.0 < lengths.mean().item() < 47.0
    w = out["winners"]
    assert 0.48 < (w == 1).float().mean().item() < 0.57
    assert int(lengths.sum().item()) == out["board_x"].shape[0] == out["v_y"].shape[0]
    # sharding invariance: the second half generated as its own shard gives the same games
    a = BatchedGreedyGenerator(eng, seed=SEED).generate(512)
    b1 = BatchedGreedyGenerator(eng, seed=SEED, rank=0, world=2).generate(256)
    b2 = BatchedGreedyGenerator(eng, seed=SEED, rank=1, world=2).generate(256)
    assert torch.equal(a["lengths"], torch.cat([b1["lengths"], b2["lengths"]]))
    assert torch.equal(a["state"], torch.cat([b1["state"], b2["state"]], dim=1))


def test_dropin_generate_play_matches_batched_records(eng):
    """GreedyDataGenerator().generate_play() (data_generators.py:23-80 surface): [(Board, pi)], reward — and
    utils.convert_to_train_data on it reproduces the batched generator's board_x / pi_y / v_y for the same games."""
    from chinesecheckersagent_b200 import utils
    from chinesecheckersagent_b200.data_generators import BatchedGreedyGenerator, GreedyDataGenerator
    gen = GreedyDataGenerator(engine=eng, batch=8, seed=SEED)
    games = [gen.generate_play() for _ in range(8)]
    ref = BatchedGreedyGenerator(eng, seed=SEED).generate(8)
    lengths = ref["lengths"].cpu().numpy()
    assert [len(h) for h, _ in games] == list(lengths)
    winners = ref["winners"].cpu().numpy()
    assert [r for _, r in games] == [{0: 0, 1: 1, 2: -1}[int(w)] for w in winners]
    bx, py, vy = utils.convert_to_train_data(games[:3])
    m = int(lengths[:3].sum())
    assert np.array_equal(np.stack(bx).astype(np.uint8), ref["board_x"][:m].cpu().numpy())
    assert np.allclose(np.stack(py), ref["pi_y"][:m].cpu().numpy().astype(np.float64), atol=1e-7)
    assert list(vy) == list(ref["v_y"][:m].cpu().numpy())
    board, pi = games[0][0][0]
    assert board.get_valid_moves(1) and abs(pi.sum() - 1.0) < 1e-6
