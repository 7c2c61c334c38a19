"""GPU parity tests for the env kernels, through the C-ABI (ctypes -> libccx.so).  Bit-exact bar."""
import numpy as np
import pytest
import torch

import oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from chinesecheckersagent_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


def env_from(eng, st):
    from chinesecheckersagent_b200.engine import BatchedEnv
    return BatchedEnv(st.shape[1], engine=eng, state=st)


def u64(t):
    return t.cpu().numpy().view(np.uint64)


def canonical_masks(ref_moves, ref_nmoves):
    m = np.zeros((6, ref_moves.shape[0]), dtype=np.uint64)
    for i in range(ref_moves.shape[0]):
        for cid in range(6):
            v = 0
            for k in range(ref_nmoves[i, cid]):
                v |= 1 << int(ref_moves[i, cid, k])
            m[cid, i] = v
    return m


def test_reset_start_matches_board_ctor(eng):
    from chinesecheckersagent_b200.engine import BatchedEnv
    env = BatchedEnv(1000, engine=eng)
    assert np.array_equal(env.numpy_state(), orc.start_states(1000))


def test_movegen_golden(eng, env_golden):
    env = env_from(eng, env_golden["state"])
    got = u64(env.movegen())
    assert np.array_equal(got, canonical_masks(env_golden["ref_moves"], env_golden["ref_nmoves"]))


def test_movegen_random_boards_vs_oracle(eng):
    st = orc.random_states(50000, seed=11)
    got = u64(env_from(eng, st).movegen())
    assert np.array_equal(got, orc.movegen(st))


def test_apply_and_win_golden(eng, env_golden):
    g = env_golden
    sel = g["chosen"][:, 0] != 255
    st = np.ascontiguousarray(g["state"][:, sel])
    env = env_from(eng, st)
    frm = torch.from_numpy(np.ascontiguousarray(g["chosen"][sel, 0])).cuda()
    to = torch.from_numpy(np.ascontiguousarray(g["chosen"][sel, 1])).cuda()
    winner = env.apply(frm, to).cpu().numpy()
    assert np.array_equal(winner, g["winner"][sel])
    assert np.array_equal(env.numpy_state()[:7], g["succ"][:7, sel])


def test_info_golden(eng, env_golden):
    g = env_golden
    info = env_from(eng, g["state"]).info().cpu().numpy()
    assert np.array_equal(info[:, 0], g["check_win"])
    assert np.array_equal(info[:, 1:3], g["progress"])
    assert np.array_equal(info[:, 3:5], g["fwd_dist"])


def test_greedy_candidates_golden(eng, env_golden):
    st = env_golden["state"]
    got = u64(env_from(eng, st).greedy_candidates())
    assert np.array_equal(got, orc.greedy_candidates(st))


@pytest.mark.parametrize("dtype_name", ["u8", "bf16", "f32"])
def test_encode_golden(eng, env_golden, dtype_name):
    from chinesecheckersagent_b200 import config as C
    dt = {"u8": C.DTYPE_U8, "bf16": C.DTYPE_BF16, "f32": C.DTYPE_F32}[dtype_name]
    planes = env_from(eng, env_golden["state"]).encode(dt)
    got = planes.float().cpu().numpy().astype(np.uint8)
    assert got.shape == env_golden["planes"].shape
    assert np.array_equal(got, env_golden["planes"])


@pytest.mark.parametrize("n", [1, 31, 129, 1000])
def test_encode_ragged_sizes(eng, n):
    from chinesecheckersagent_b200 import config as C
    st, _, _ = orc.step_random(orc.start_states(n), 99, 0, 9)
    for dt in (C.DTYPE_U8, C.DTYPE_BF16, C.DTYPE_F32):
        got = env_from(eng, st).encode(dt).float().cpu().numpy().astype(np.uint8)
        assert np.array_equal(got, orc.encode(st))


def test_step_random_trace_bit_exact(eng):
    """cfg 2 parity sample: first 1024 games, every ply's state, masks, chosen move, winner."""
    from chinesecheckersagent_b200.engine import BatchedEnv
    n, plies, tg = 4096, 64, 1024
    env = BatchedEnv(n, engine=eng, seed=0x5EED2026, game_id0=7)
    trace = u64(env.step_random(plies, trace_games=tg))
    ost, owins, otrace = orc.step_random(orc.start_states(n), 0x5EED2026, 0, plies, game_id0=7, trace_games=tg, nthreads=8)
    assert np.array_equal(trace, otrace)
    assert np.array_equal(env.numpy_state()[:5], ost[:5])
    assert np.array_equal(env.wins.cpu().numpy().view(np.uint64), owins)


def test_step_random_chunking_invariant(eng):
    """256 plies in one launch == 4 launches of 64 (counter RNG keyed by absolute step)."""
    from chinesecheckersagent_b200.engine import BatchedEnv
    a = BatchedEnv(2048, engine=eng)
    a.step_random(256)
    b = BatchedEnv(2048, engine=eng)
    for _ in range(4):
        b.step_random(64)
    assert torch.equal(a.state[:5], b.state[:5])


def test_step_random_full_size_properties(eng):
    """BASELINE cfg 2 size (65,536 games): invariants that do not need the oracle at full size, plus a
    bit-exact spot check of a strided sample against the oracle."""
    from chinesecheckersagent_b200.engine import BatchedEnv
    n, plies = 65536, 256
    env = BatchedEnv(n, engine=eng)
    env.step_random(plies)
    st = env.numpy_state()
    occ1, occ2 = st[0], st[1]
    valid = np.uint64(0x007F7F7F7F7F7F7F)
    assert np.all((occ1 & occ2) == 0) and np.all(((occ1 | occ2) & ~valid) == 0)
    pop = np.array([bin(int(x)).count("1") for x in occ1[:4096]])
    assert np.all(pop == 6)
    # cells words agree with occupancy
    for w, occ in ((st[2], occ1), (st[3], occ2)):
        rebuilt = np.zeros(n, dtype=np.uint64)
        for k in range(6):
            rebuilt |= np.uint64(1) << ((w >> np.uint64(8 * k)) & np.uint64(0xFF))
        assert np.array_equal(rebuilt, occ)
    # game ids are global: a shard of the same ids computed alone gives identical states
    shard = BatchedEnv(1024, engine=eng, game_id0=30000)
    shard.step_random(plies)
    assert torch.equal(shard.state[:5], env.state[:5, 30000:31024])
    ost, _, _ = orc.step_random(orc.start_states(256), env.seed, 0, plies, game_id0=30000, nthreads=8)
    assert np.array_equal(st[:5, 30000:30256], ost[:5])


def test_play_greedy_matches_reference_games(eng, greedy_golden):
    from chinesecheckersagent_b200.engine import BatchedEnv
    g = greedy_golden
    n = g["status"].shape[0]
    env = BatchedEnv(n, engine=eng, seed=int(g["seed"]))
    counters = env.play_greedy().cpu().numpy()
    st = env.numpy_state()
    assert np.array_equal(st[:7], g["final"][:7])
    assert counters[0] == g["plies"].sum()
    assert counters[1] == (g["status"] == 1).sum() and counters[2] == (g["status"] == 2).sum()
    assert counters[3] == (g["status"] == 3).sum()


def test_play_greedy_large_vs_oracle_and_stats(eng):
    from chinesecheckersagent_b200.engine import BatchedEnv
    n = 20000
    env = BatchedEnv(n, engine=eng, seed=4242)
    c = env.play_greedy().cpu().numpy()
    ost = orc.play_greedy(orc.start_states(n), 4242, nthreads=8)
    assert np.array_equal(env.numpy_state()[:7], ost[:7])
    # reference statistics (config.py:77 AVERAGE_TOTAL_MOVE = 43; SURVEY §6: 43.3 plies, ~1 % stopped)
    assert 42.0 < c[0] / n < 45.0
    assert c[3] / n < 0.03


def test_randomised_reset_is_a_valid_placement(eng):
    from chinesecheckersagent_b200.engine import BatchedEnv
    env = BatchedEnv(5000, engine=eng, randomised=True, seed=5)
    st = env.numpy_state()
    assert np.all((st[0] & st[1]) == 0)
    both = st[0] | st[1]
    assert all(bin(int(x)).count("1") == 12 for x in both)
    assert len(np.unique(both)) > 4900            # essentially all distinct
    # the kernels accept them: masks agree with the oracle
    assert np.array_equal(u64(env.movegen()), orc.movegen(st))


def test_randomised_reset_is_uniform_per_checker(eng):
    """board.py:69 `np.random.choice(49, 12, replace=False)`: every one of the 12 draws (player, checker id) is marginally
    uniform over the 49 cells, and an ordered pair of draws is uniform over the 49 * 48 ordered cell pairs.  Chi-square on
    196,608 boards at p = 1e-5 (48 degrees of freedom: 103.4; 2,351: 2,652)."""
    from chinesecheckersagent_b200.engine import BatchedEnv
    n = 196608
    st = BatchedEnv(n, engine=eng, randomised=True, seed=2026).numpy_state()
    cells = np.zeros((12, n), dtype=np.int64)
    for pl in (0, 1):
        for i in range(6):
            c = ((st[2 + pl] >> np.uint64(8 * i)) & np.uint64(0xFF)).astype(np.int64)
            cells[pl * 6 + i] = (c >> 3) * 7 + (c & 7)
    assert cells.min() >= 0 and cells.max() < 49
    for k in range(12):
        counts = np.bincount(cells[k], minlength=49)
        chi2 = float(((counts - n / 49) ** 2 / (n / 49)).sum())
        assert chi2 < 103.4, (k, chi2)
    for a, b in ((0, 1), (5, 6), (3, 11)):
        pair = np.bincount(cells[a] * 49 + cells[b], minlength=49 * 49).reshape(49, 49)
        assert np.all(np.diag(pair) == 0)                              # without replacement
        off = pair[~np.eye(49, dtype=bool)]
        exp = n / (49 * 48)
        assert float(((off - exp) ** 2 / exp).sum()) < 2652, (a, b)


def test_host_buffer_variants(eng, env_golden):
    from chinesecheckersagent_b200.engine import HostEnv
    from chinesecheckersagent_b200 import config as C
    h = HostEnv(engine=eng)
    st = np.ascontiguousarray(env_golden["state"])
    assert np.array_equal(h.movegen(st), canonical_masks(env_golden["ref_moves"], env_golden["ref_nmoves"]))
    assert np.array_equal(h.encode(st, C.DTYPE_U8), env_golden["planes"])
    s0 = orc.start_states(512)
    wins = h.step_random(s0, 100, seed=3, step0=0, game_id0=0)
    ost, owins, _ = orc.step_random(orc.start_states(512), 3, 0, 100)
    assert np.array_equal(s0[:5], ost[:5]) and np.array_equal(wins, owins)


def test_empty_batch_is_ok(eng):
    from chinesecheckersagent_b200 import _lib
    assert eng.L.ccx_movegen(eng.h, 0, None, None) == 0
    assert eng.L.ccx_step_random(eng.h, 0, None, 0, 0, 0, 5, None, None, 0) == 0
