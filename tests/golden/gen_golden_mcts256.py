"""Second MCTS fixture (SURVEY.md §8d cfg 4: "per-edge N identical to the shimmed reference on >= 256 roots"): 256 seeded
roots x the two stub-evaluator paths — AiPlayer.decide_move (player.py:157-158: unexpanded root, sum N = 174) and make_move
without noise (selfplay.py:114-127: root expanded first, sum N = 175) — searched by the UNMODIFIED reference MCTS.py with
the harness of gen_golden_mcts.py (first-choice tie-break, canonical edge order).  Roots: 160 cfg-4 roots (start position
advanced by 6 random plies), 64 `Board(randomised=True)` placements (board.py:61-85), 32 mid-game positions.
Build container only; ~1 min on 8 cores.  Writes tests/golden/mcts_golden_256.npz."""
import multiprocessing as mp
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)


def make_roots():
    import gen_golden_mcts as G
    from gen_golden import pack_board, random_move, SEED
    Board = G.Board
    rnd = random.Random(SEED + 256)
    roots = []
    for _ in range(160):
        b, player = Board(), 1
        for _ in range(6):
            b.place(player, *random_move(b, player, rnd)); player = 3 - player
        roots.append(pack_board(b, player - 1))
    np.random.seed(SEED % (2 ** 31))
    for i in range(64):
        b = Board(randomised=True)                                   # np.random.choice(49, 12, replace=False), board.py:69
        roots.append(pack_board(b, i & 1))
    for _ in range(32):
        b, player = Board(), 1
        for _ in range(rnd.randint(15, 60)):
            b.place(player, *random_move(b, player, rnd)); player = 3 - player
        roots.append(pack_board(b, player - 1))
    return np.ascontiguousarray(np.array(roots, dtype=np.uint64).T)


def board_from_words(G, w):
    """reference Board carrying the packed position (oracle helper of SURVEY §8c: overwrite board / checkers_pos / checkers_id)"""
    b = G.Board()
    b.board[:] = 0
    for pl in (1, 2):
        cells = int(w[pl + 1])
        pos = {i: (((cells >> (8 * i)) & 0xFF) >> 3, ((cells >> (8 * i)) & 0xFF) & 7) for i in range(6)}
        b.checkers_pos[pl] = pos
        b.checkers_id[pl] = {p: i for i, p in pos.items()}
        for p in pos.values():
            b.board[p[0], p[1], 0] = pl
    b.hist_moves.clear()
    return b


def search_job(args):
    ri, words, pre_expand = args
    import gen_golden_mcts as G
    G.install_harness()
    b = board_from_words(G, words)
    player = int((int(words[4]) >> 48) & 1) + 1
    v, pi, q, nn, ne = G.run_search(b, player, (0, pre_expand, 0, 1.0, 175), None)
    return ri, pre_expand, v, pi, q, nn


def main():
    roots = make_roots()
    n = roots.shape[1]
    # history-free roots: the stub evaluator ignores the planes, so only occupancy / ids / side to move matter; keep ply 0 metadata
    roots[4] = (roots[4] & np.uint64(0x0001000000000000)) | np.uint64(0x00000000FFFFFFFF)
    roots[5:7] = np.uint64(0xFFFFFFFFFFFFFFFF)
    roots[7] = 0
    jobs = [(ri, roots[:, ri].copy(), pe) for pe in (0, 1) for ri in range(n)]
    with mp.Pool(os.cpu_count()) as pool:
        res = pool.map(search_job, jobs, chunksize=8)
    out = {"roots": roots}
    for pe in (0, 1):
        V = np.zeros((n, 294), np.uint32); P = np.zeros((n, 294)); Q = np.zeros((n, 294)); NN = np.zeros(n, np.int32)
        for ri, p, v, pi, q, nn in res:
            if p == pe:
                V[ri], P[ri], Q[ri], NN[ri] = v, pi, q, nn
        out["visits%d" % pe], out["pi%d" % pe], out["q%d" % pe], out["nodes%d" % pe] = V, P, Q, NN
        print("pre_expand", pe, "sum N", set(V.sum(1).tolist()), "mean nodes", NN.mean())
    np.savez_compressed(os.path.join(HERE, "mcts_golden_256.npz"), **out)


if __name__ == "__main__":
    main()
