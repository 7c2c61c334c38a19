"""Pins the CPU oracle (oracle/ccx_oracle.c) against fixtures produced by the unmodified reference.

CPU only.  Every check is bit-exact (integer / index work)."""
import numpy as np

import oracle as orc

# Random123 known-answer vectors for Philox4x32-10 (kat_vectors: ctr, key -> out)
PHILOX_KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def test_philox_known_answers():
    for ctr, key, out in PHILOX_KAT:
        got = orc.philox(key[0], key[1], *ctr)
        assert tuple(int(x) for x in got) == out


def test_start_position_known_answer():
    # P1's start move list, policy idx {70,86,127,143,168,176,225,233,282,290}
    st = orc.start_states(1)
    out, cnt = orc.movelist(st)
    idx = set()
    for cid in range(6):
        for k in range(cnt[0, cid]):
            r, c = orc.cell_rc(out[0, cid, k])
            idx.add(cid * 49 + r * 7 + c)
    assert idx == {70, 86, 127, 143, 168, 176, 225, 233, 282, 290}
    assert list(cnt[0]) == [0, 2, 2, 2, 2, 2]


def test_movelists_reference_order(env_golden):
    out, cnt = orc.movelist(env_golden["state"])
    assert np.array_equal(cnt, env_golden["ref_nmoves"])
    assert np.array_equal(out, env_golden["ref_moves"])


def test_place_successor_and_winner(env_golden):
    g = env_golden
    sel = g["chosen"][:, 0] != 255
    st = np.ascontiguousarray(g["state"][:, sel])
    succ, winner = orc.apply(st, g["chosen"][sel, 0], g["chosen"][sel, 1])
    assert np.array_equal(winner, g["winner"][sel])
    assert np.array_equal(succ[:7], g["succ"][:7, sel])
    assert int((winner > 0).sum()) >= 20


def test_win_progress_distance(env_golden):
    g = env_golden
    info = orc.info(g["state"])
    assert np.array_equal(info[:, 0], g["check_win"])
    assert np.array_equal(info[:, 1:3], g["progress"])
    assert np.array_equal(info[:, 3:5], g["fwd_dist"])
    assert set(np.unique(g["check_win"])) == {0, 1, 2}


def test_to_model_input(env_golden):
    planes = orc.encode(env_golden["state"])
    assert np.array_equal(planes, env_golden["planes"])


def test_greedy_candidates(env_golden):
    g = env_golden
    out, cnt = orc.greedy_list(g["state"])
    assert np.array_equal(cnt, g["n_greedy"])
    assert np.array_equal(out, g["greedy"])


def test_greedy_games_match_reference_game_loop(greedy_golden):
    g = greedy_golden
    n = g["status"].shape[0]
    st = orc.play_greedy(orc.start_states(n), int(g["seed"]), max_plies=100000)
    meta = st[4]
    status = (meta >> np.uint64(56)).astype(np.uint8)
    plies = ((meta >> np.uint64(32)) & np.uint64(0xFFFF)).astype(np.int32)
    assert np.array_equal(status, g["status"])
    assert np.array_equal(plies, g["plies"])
    assert np.array_equal(st[:7], g["final"][:7])
    assert (status == 3).sum() >= 1     # at least one repetition stop is pinned


def test_threads_do_not_change_results():
    st0 = orc.start_states(64)
    a, wa, _ = orc.step_random(st0, 123, 0, 40, nthreads=1)
    b, wb, _ = orc.step_random(st0, 123, 0, 40, nthreads=4)
    assert np.array_equal(a, b) and np.array_equal(wa, wb)
