"""MCTS golden fixture: runs the UNMODIFIED reference MCTS.py on seeded roots (build container only).

Harness-level choices (documented in DESIGN.md): `MCTS.random` is replaced by a shim
whose choice() returns the first maximal edge, and Board.get_valid_moves is wrapped so that each
checker's destination list is sorted by r*7+c (canonical edge order).  Evaluators return float64."""
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, HERE)
import refshim  # noqa: E402
import oracle as orc  # noqa: E402
from gen_golden import pack_board, random_move, SEED  # noqa: E402

R = refshim.load()
M = R.MCTS
Board = R.board.Board

CONFIGS = [  # (evaluator, pre_expand, use_noise, tau, num_itr)
    (0, 0, 0, 1.0, 175),
    (0, 1, 0, 1.0, 175),
    (1, 0, 0, 1.0, 175),
    (1, 1, 1, 1.0, 175),
    (1, 0, 0, 0.01, 175),
]
NOISE_STRIDE = 128


# ---- engine-side test evaluators, numpy restatement (spec in oracle/ccx_oracle_mcts.c) ------------
def splitmix64(z):
    M64 = (1 << 64) - 1
    z = (z + 0x9E3779B97F4A7C15) & M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    return z ^ (z >> 31)


COEF = [splitmix64(k + 1) for k in range(343)]


def philox_vec(k0, k1, c0):
    """Philox4x32-10 over a vector of c0 counters with c1 = 7, c2 = c3 = 0; returns word 0."""
    c0 = c0.astype(np.uint64)
    c1 = np.full_like(c0, 7)
    c2 = np.zeros_like(c0)
    c3 = np.zeros_like(c0)
    m32 = np.uint64(0xFFFFFFFF)
    k0 = np.uint64(k0); k1 = np.uint64(k1)
    for _ in range(10):
        p0 = np.uint64(0xD2511F53) * c0
        p1 = np.uint64(0xCD9E8D57) * c2
        c0, c1, c2, c3 = ((p1 >> np.uint64(32)) ^ c1 ^ k0) & m32, p1 & m32, ((p0 >> np.uint64(32)) ^ c3 ^ k1) & m32, p0 & m32
        k0 = (k0 + np.uint64(0x9E3779B9)) & m32
        k1 = (k1 + np.uint64(0xBB67AE85)) & m32
    return c0


class UniformModel:
    version = 0

    def predict(self, x):
        return np.full(294, 1 / 294.), 0.0          # Python float: W, Q stay float64


class HashModel:
    version = 0

    def predict(self, x):
        flat = x.astype(np.uint8).reshape(-1)
        h = 0
        for k in np.nonzero(flat)[0]:
            h = (h + COEF[k] * int(flat[k])) & ((1 << 64) - 1)
        w = philox_vec(h & 0xFFFFFFFF, h >> 32, np.arange(295))
        p = w[:294].astype(np.float64) / 4294967296.0
        v = float(w[294]) / 2147483648.0 - 1.0
        return p, v


def install_harness():
    M.random = refshim.FirstChoice
    if not getattr(Board, "_ccx_canonical", False):
        orig = Board.get_valid_moves

        def canonical(self, cur_player):
            vm = orig(self, cur_player)
            return {k: sorted(v, key=lambda rc: rc[0] * 7 + rc[1]) for k, v in vm.items()}
        Board.get_valid_moves = canonical
        Board._ccx_canonical = True


def make_roots():
    rnd = random.Random(SEED + 1)
    roots = []
    for i in range(32):                      # cfg 4 roots: start advanced by 6 random plies
        b = Board(); player = 1
        for _ in range(6):
            b.place(player, *random_move(b, player, rnd)); player = 3 - player
        b._ccx_plies = 6
        roots.append((b, player))
    for i in range(8):                       # mid-game
        b = Board(); player = 1
        n = rnd.randint(15, 40)
        for _ in range(n):
            b.place(player, *random_move(b, player, rnd)); player = 3 - player
        b._ccx_plies = n
        roots.append((b, player))
    GP = R.player.GreedyPlayer
    while len(roots) < 48:                   # late greedy positions: wins inside the search horizon
        b = Board(); player = 1; hist = []
        import copy
        for ply in range(200):
            hist.append((copy.deepcopy(b), player, ply))
            s, e = rnd.choice(GP(player).decide_move(b, training=True))
            if b.place(player, R.board_utils.human_coord_to_np_index(s), R.board_utils.human_coord_to_np_index(e)):
                break
            player = 3 - player
        bb, pl, ply = hist[-rnd.randint(1, 3)]
        bb._ccx_plies = ply
        roots.append((bb, pl))
    return roots


def run_search(board, player, cfg, noise):
    import copy
    evaluator, pre_expand, use_noise, tau, num_itr = cfg
    model = HashModel() if evaluator else UniformModel()
    root = M.Node(copy.deepcopy(board), player)
    tree = M.MCTS(root, model, cpuct=3.5, num_itr=num_itr, tree_tau=tau)
    if pre_expand:                                            # selfplay.py:114-124
        tree.expandAndBackUp(tree.root, breadcrumbs=[])
        if use_noise:
            for i in range(len(tree.root.edges)):
                tree.root.edges[i].stats['P'] *= (1. - 0.25)
                tree.root.edges[i].stats['P'] += 0.25 * noise[i]
    np.random.seed(1)
    pi, _ = tree.search()
    visits = np.zeros(294, dtype=np.uint32)
    q = np.zeros(294, dtype=np.float64)
    for e in root.edges:
        cid = root.state.checkers_id[player][e.fromPos]
        idx = R.utils.encode_checker_index(cid, e.toPos)
        visits[idx] = e.stats['N']
        q[idx] = e.stats['Q']
    n_nodes = 0
    stack = [root]
    while stack:
        nd = stack.pop(); n_nodes += 1
        stack.extend(e.outNode for e in nd.edges)
    return visits, np.array(pi, dtype=np.float64), q, n_nodes, len(root.edges)


def main():
    install_harness()
    roots = make_roots()
    rng = np.random.default_rng(SEED)
    states = np.ascontiguousarray(np.array([pack_board(b, p - 1) for b, p in roots], dtype=np.uint64).T)
    noise = np.zeros((len(roots), NOISE_STRIDE), dtype=np.float64)
    out = {"roots": states, "configs": np.array(CONFIGS, dtype=np.float64), "noise": noise}
    for ci, cfg in enumerate(CONFIGS):
        V, P, Q, NN = [], [], [], []
        for ri, (b, p) in enumerate(roots):
            if cfg[2]:
                ne = sum(len(v) for v in b.get_valid_moves(p).values())
                noise[ri, :ne] = rng.dirichlet(np.ones(ne) * 0.03)          # selfplay.py:121
            v, pi, q, nn, _ = run_search(b, p, cfg, noise[ri])
            V.append(v); P.append(pi); Q.append(q); NN.append(nn)
        out["visits%d" % ci] = np.array(V); out["pi%d" % ci] = np.array(P)
        out["q%d" % ci] = np.array(Q); out["nodes%d" % ci] = np.array(NN, dtype=np.int32)
        print("cfg", ci, cfg, "sumN", out["visits%d" % ci].sum(1)[:4], "nodes mean", np.mean(NN))
    np.savez_compressed(os.path.join(HERE, "mcts_golden.npz"), **out)


if __name__ == "__main__":
    main()
