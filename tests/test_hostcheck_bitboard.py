"""CPU-side check of the kernels' bitboard SOURCE (ccx_device.cuh compiled for the host, test-only)
against the oracle and the reference fixtures.  The GPU parity tests proper are in test_gpu_*.py."""
import ctypes
import os
import sys

import numpy as np
import pytest

import oracle as orc

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "hostcheck"))
import build as hc_build  # noqa: E402


@pytest.fixture(scope="module")
def hc():
    L = ctypes.CDLL(hc_build.build())
    vp, i64, i32, u64, u32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_uint64, ctypes.c_uint32
    L.hc_movegen.argtypes = [vp, i64, vp]
    L.hc_greedy.argtypes = [vp, i64, vp]
    L.hc_movegen_rays.argtypes = [vp, i64, vp]
    L.hc_movegen_tri.argtypes = [vp, i64, vp, i32]
    L.hc_apply.argtypes = [vp, i64, vp, vp, vp]
    L.hc_step_random.argtypes = [vp, i64, i64, u64, u32, i32, vp, vp, i64]
    L.hc_play_greedy.argtypes = [vp, i64, i64, u64, i32]
    return L


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def canonical_masks(ref_moves, ref_nmoves):
    m = np.zeros((6, ref_moves.shape[0]), dtype=np.uint64)
    for i in range(ref_moves.shape[0]):
        for cid in range(6):
            v = 0
            for k in range(ref_nmoves[i, cid]):
                v |= 1 << int(ref_moves[i, cid, k])
            m[cid, i] = v
    return m


def test_movegen_vs_reference_fixture(hc, env_golden):
    st = np.ascontiguousarray(env_golden["state"])
    n = st.shape[1]
    masks = np.zeros((6, n), dtype=np.uint64)
    hc.hc_movegen(P(st), n, P(masks))
    assert np.array_equal(masks, canonical_masks(env_golden["ref_moves"], env_golden["ref_nmoves"]))


def test_movegen_vs_oracle_random_boards(hc):
    st = orc.random_states(20000, seed=7)
    masks = np.zeros((6, st.shape[1]), dtype=np.uint64)
    hc.hc_movegen(P(st), st.shape[1], P(masks))
    assert np.array_equal(masks, orc.movegen(st))


def test_ray_movegen_vs_reference_and_oracle(hc, env_golden):
    """The table-driven ray formulation (movegen_rays) == the reference on the fixture and == the oracle on
    40,000 random boards and on positions from long random games."""
    st = np.ascontiguousarray(env_golden["state"])
    masks = np.zeros((6, st.shape[1]), dtype=np.uint64)
    hc.hc_movegen_rays(P(st), st.shape[1], P(masks))
    assert np.array_equal(masks, canonical_masks(env_golden["ref_moves"], env_golden["ref_nmoves"]))
    st = orc.random_states(40000, seed=21)
    masks = np.zeros((6, st.shape[1]), dtype=np.uint64)
    hc.hc_movegen_rays(P(st), st.shape[1], P(masks))
    assert np.array_equal(masks, orc.movegen(st))
    st, _, _ = orc.step_random(orc.start_states(4000), 5, 0, 150, nthreads=8)
    masks = np.zeros((6, st.shape[1]), dtype=np.uint64)
    hc.hc_movegen_rays(P(st), st.shape[1], P(masks))
    assert np.array_equal(masks, orc.movegen(st))


def test_apply_vs_reference_fixture(hc, env_golden):
    g = env_golden
    sel = g["chosen"][:, 0] != 255
    st = np.ascontiguousarray(g["state"][:, sel])
    frm = np.ascontiguousarray(g["chosen"][sel, 0]); to = np.ascontiguousarray(g["chosen"][sel, 1])
    winner = np.zeros(st.shape[1], dtype=np.uint8)
    hc.hc_apply(P(st), st.shape[1], P(frm), P(to), P(winner))
    assert np.array_equal(winner, g["winner"][sel])
    assert np.array_equal(st[:7], g["succ"][:7, sel])


def test_greedy_vs_reference_fixture(hc, env_golden):
    g = env_golden
    st = np.ascontiguousarray(g["state"])
    n = st.shape[1]
    masks = np.zeros((6, n), dtype=np.uint64)
    hc.hc_greedy(P(st), n, P(masks))
    assert np.array_equal(masks, orc.greedy_candidates(st))
    # and directly against the reference's candidate lists
    for i in range(0, n, 7):
        want = set()
        for k in range(g["n_greedy"][i]):
            want.add((int(g["greedy"][i, k, 0]), int(g["greedy"][i, k, 1])))
        cells = int(st[2 + ((int(st[4, i]) >> 48) & 1), i])
        got = set()
        for cid in range(6):
            m = int(masks[cid, i])
            for b in range(64):
                if (m >> b) & 1:
                    got.add(((cells >> (8 * cid)) & 0xFF, b))
        assert got == want


def test_step_random_trace_vs_oracle(hc):
    n, plies, tg = 96, 120, 96
    st = orc.start_states(n)
    wins = np.zeros(2, dtype=np.uint64)
    trace = np.zeros((plies, tg, 12), dtype=np.uint64)
    hc.hc_step_random(P(st), n, 1000, 0x5EED2026, 5, plies, P(wins), P(trace), tg)
    ost, owins, otrace = orc.step_random(orc.start_states(n), 0x5EED2026, 5, plies, game_id0=1000, trace_games=tg)
    assert np.array_equal(trace, otrace)
    assert np.array_equal(st[:5], ost[:5])
    assert np.array_equal(wins, owins)


def test_play_greedy_vs_reference_games(hc, greedy_golden):
    g = greedy_golden
    n = g["status"].shape[0]
    st = orc.start_states(n)
    hc.hc_play_greedy(P(st), n, 0, int(g["seed"]), 100000)
    assert np.array_equal(st[:7], g["final"][:7])


@pytest.mark.parametrize("use_lut2", [0, 1])
def test_three_layout_and_occupancy_major_expansions_vs_oracle(hc, env_golden, use_lut2):
    """expand_cell_tri (row / column-major / diagonal-major occupancy copies, pre-scattered 64-bit answers) and expand_cell_lut2
    (occupancy-major byte table) give the reference's move lists on the fixture positions and the oracle's on 20,000 random
    boards and 20,000 positions reached by random play."""
    st = np.ascontiguousarray(env_golden["state"])
    masks = np.zeros((6, st.shape[1]), dtype=np.uint64)
    hc.hc_movegen_tri(P(st), st.shape[1], P(masks), use_lut2)
    assert np.array_equal(masks, canonical_masks(env_golden["ref_moves"], env_golden["ref_nmoves"]))
    for st in (orc.random_states(20000, seed=11), orc.step_random(orc.start_states(20000), 5, 0, 37, nthreads=8)[0]):
        masks = np.zeros((6, st.shape[1]), dtype=np.uint64)
        hc.hc_movegen_tri(P(st), st.shape[1], P(masks), use_lut2)
        assert np.array_equal(masks, orc.movegen(st))
