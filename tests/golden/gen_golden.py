"""Generate golden fixtures by running the UNMODIFIED reference (/root/reference) as the source of truth.

Run in the build container only:  python tests/golden/gen_golden.py [env|greedy|mcts|all]
The outputs (*.npz next to this file) are committed; the GPU box never runs this script.

What each fixture pins:
  env_golden.npz     Board.get_valid_moves (reference order), Board.place successor + winner,
                     Board.check_win, player_progress, player_forward_distance, utils.to_model_input,
                     GreedyPlayer.decide_move(training=True) on ~4.5k positions reached by random play,
                     Board(randomised=True) starts and greedy play, plus hand-made win boards.
  greedy_games.npz   game.Game.start() outcomes (winner / repetition stop, ply count, final position)
                     with the uniform pick among filtered_best_moves driven by the engine's Philox rule.
  mcts_golden.npz    MCTS.search visit counts / pi with a uniform-prior float64 stub evaluator and
                     first-maximum tie-break, AiPlayer path (unexpanded root) and make_move path.
"""
import contextlib
import io
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import refshim  # noqa: E402
import oracle as orc  # noqa: E402  (only its numpy packing helpers are used here)

R = refshim.load()
Board = R.board.Board

SEED = 0x5EED2026


# ----------------------------------------------------------------------------------------------
# Philox4x32-10 in pure Python (engine RNG spec; independent of the C/CUDA implementations)

def philox(k0, k1, c0, c1, c2, c3):
    M = 0xFFFFFFFF
    for _ in range(10):
        p0 = 0xD2511F53 * c0
        p1 = 0xCD9E8D57 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & M, p1 & M, ((p0 >> 32) ^ c3 ^ k1) & M, p0 & M
        k0 = (k0 + 0x9E3779B9) & M
        k1 = (k1 + 0xBB67AE85) & M
    return c0, c1, c2, c3


def mulhi(a, b):
    return (a * b) >> 32


# ----------------------------------------------------------------------------------------------

def pack_board(board, to_move, status=0):
    """Reference Board -> packed uint64[8] (include/ccx.h layout)."""
    p1 = [board.checkers_pos[1][i] for i in range(6)]
    p2 = [board.checkers_pos[2][i] for i in range(6)]
    hm = list(board.hist_moves)
    last = [(hm[-1 - k][0], hm[-1 - k][1]) for k in range(min(2, len(hm)))]
    dests = [hm[-1 - k][1] for k in range(len(hm))]
    ply = getattr(board, "_ccx_plies", len(hm))
    return orc.pack_state(p1, p2, to_move=to_move, ply=ply, last_moves=last, hist_dests=dests, status=status)


class Recorder:
    def __init__(self):
        self.rows = {k: [] for k in ("state", "ref_moves", "ref_nmoves", "chosen", "winner", "succ", "planes",
                                     "progress", "check_win", "fwd_dist", "greedy", "n_greedy")}

    def record(self, board, player, move=None):
        """Record everything about `board` with `player` to move; if `move` is given apply it."""
        to_move = player - 1
        rows = self.rows
        rows["state"].append(pack_board(board, to_move))
        vm = board.get_valid_moves(player)
        mv = np.full((6, 24), -1, dtype=np.int8)
        nm = np.zeros(6, dtype=np.int8)
        for cid in range(6):
            dests = vm[board.checkers_pos[player][cid]]
            nm[cid] = len(dests)
            for k, d in enumerate(dests):
                mv[cid, k] = orc.cell(*d)
        rows["ref_moves"].append(mv)
        rows["ref_nmoves"].append(nm)
        rows["planes"].append(R.utils.to_model_input(board, player).astype(np.uint8))
        rows["progress"].append([board.player_progress(1), board.player_progress(2)])
        rows["check_win"].append(board.check_win())
        rows["fwd_dist"].append([board.player_forward_distance(1), board.player_forward_distance(2)])
        g = np.full((32, 2), -1, dtype=np.int16)
        ng = 0
        if any(len(v) for v in vm.values()):
            cands = R.player.GreedyPlayer(player).decide_move(board, training=True)
            ng = len(cands)
            for k, (s, e) in enumerate(cands):
                g[k] = (orc.cell(*R.board_utils.human_coord_to_np_index(s)),
                        orc.cell(*R.board_utils.human_coord_to_np_index(e)))
        rows["greedy"].append(g)
        rows["n_greedy"].append(ng)
        if move is None:
            rows["chosen"].append([255, 255])
            rows["winner"].append(0)
            rows["succ"].append(np.zeros(8, dtype=np.uint64))
            return 0
        frm, to = move
        winner = board.place(player, frm, to)
        board._ccx_plies = getattr(board, "_ccx_plies", len(board.hist_moves) - 1) + 1
        rows["chosen"].append([orc.cell(*frm), orc.cell(*to)])
        rows["winner"].append(winner)
        rows["succ"].append(pack_board(board, 1 - to_move))
        return winner

    def arrays(self):
        out = {}
        for k, v in self.rows.items():
            a = np.array(v)
            if k in ("state", "succ"):
                a = np.ascontiguousarray(a.astype(np.uint64).T)
            out[k] = a
        out["chosen"] = out["chosen"].astype(np.uint8)
        out["winner"] = out["winner"].astype(np.uint8)
        out["progress"] = out["progress"].astype(np.uint8)
        out["check_win"] = out["check_win"].astype(np.uint8)
        out["fwd_dist"] = out["fwd_dist"].astype(np.int16)
        out["n_greedy"] = out["n_greedy"].astype(np.int16)
        return out


def random_move(board, player, rnd):
    vm = board.get_valid_moves(player)
    starts = [s for s in vm if vm[s]]
    if not starts:
        return None
    s = rnd.choice(starts)
    return s, rnd.choice(vm[s])


def gen_env():
    rnd = random.Random(SEED)
    np.random.seed(SEED & 0x7FFFFFFF)
    rec = Recorder()
    # (a) random walks from the start position
    for _ in range(24):
        b = Board()
        player = 1
        for _ in range(80):
            mv = random_move(b, player, rnd)
            if rec.record(b, player, mv):
                break
            player = 3 - player
    # (b) Board(randomised=True) + short random walks (history planes start empty, board.py:61-85)
    for _ in range(120):
        b = Board(randomised=True)
        player = 1
        for _ in range(6):
            mv = random_move(b, player, rnd)
            if mv is None or rec.record(b, player, mv):
                break
            player = 3 - player
    # (c) greedy play to the end (wins, progress counters, long jump chains in dense middle games)
    for _ in range(30):
        b = Board()
        player = 1
        for _ in range(200):
            cands = R.player.GreedyPlayer(player).decide_move(b, training=True)
            s, e = rnd.choice(cands)
            mv = (R.board_utils.human_coord_to_np_index(s), R.board_utils.human_coord_to_np_index(e))
            if rec.record(b, player, mv):
                rec.record(b, 3 - player, None)      # terminal position itself (check_win != 0)
                break
            player = 3 - player
    # (d) hand-made boards: both sides "won" (board.py:111 returns PLAYER_ONE), each side alone, long jumps
    t1 = [(0, 4), (0, 5), (0, 6), (1, 5), (1, 6), (2, 6)]
    t2 = [(4, 0), (5, 0), (5, 1), (6, 0), (6, 1), (6, 2)]
    mid = [(3, 3), (3, 2), (2, 2), (4, 4), (3, 4), (2, 3)]
    for p1, p2 in ((t1, t2), (t1, mid), (mid, t2), ([(6, 0), (4, 0), (3, 3), (2, 2), (6, 6), (0, 0)],
                                                   [(6, 3), (3, 0), (1, 1), (5, 5), (3, 6), (0, 3)])):
        for player in (1, 2):
            b = Board()
            b.board[:, :, 0] = 0
            b.checkers_pos = [None, {}, {}]
            b.checkers_id = [None, {}, {}]
            for pl, cells in ((1, p1), (2, p2)):
                for cid, pos in enumerate(cells):
                    b.board[pos[0], pos[1], 0] = pl
                    b.checkers_pos[pl][cid] = pos
                    b.checkers_id[pl][pos] = cid
            rec.record(b, player, None)
    arrs = rec.arrays()
    np.savez_compressed(os.path.join(HERE, "env_golden.npz"), **arrs)
    print("env_golden: %d positions, wins recorded: %d" % (arrs["state"].shape[1], int((arrs["winner"] > 0).sum())))


# ----------------------------------------------------------------------------------------------

def gen_greedy_games(n_games=300):
    """Game.start (game.py:58-100) with both GreedyPlayers picking by the engine's Philox rule."""
    GP = R.player.GreedyPlayer

    class PhiloxGreedy(GP):
        def __init__(self, player_num, gid):
            GP.__init__(self, player_num)
            self.gid = gid

        def decide_move(self, board, verbose=False, training=False, total_moves=None):
            cands = GP.decide_move(self, board, training=True)
            keyed = []
            for s, e in cands:
                frm = R.board_utils.human_coord_to_np_index(s)
                to = R.board_utils.human_coord_to_np_index(e)
                keyed.append((board.checkers_id[self.player_num][frm] * 64 + orc.cell(*to), frm, to))
            keyed.sort()
            r = philox(SEED & 0xFFFFFFFF, SEED >> 32, total_moves, 1, self.gid & 0xFFFFFFFF, self.gid >> 32)
            _, frm, to = keyed[mulhi(r[0], len(keyed))]
            return frm, to

    winners, plies, finals = [], [], []
    for gid in range(n_games):
        g = R.game.Game(p1_type="greedy", p2_type="greedy", verbose=False)
        g.player_one = PhiloxGreedy(1, gid)
        g.player_two = PhiloxGreedy(2, gid)
        g.cur_player, g.next_player = g.player_one, g.player_two
        places = [0]
        orig_place = g.board.place

        def counting_place(*a, _o=orig_place, _p=places):
            _p[0] += 1
            return _o(*a)
        g.board.place = counting_place
        with contextlib.redirect_stdout(io.StringIO()):
            w = g.start()
        g.board._ccx_plies = places[0]
        status = w if w else 3
        winners.append(status)
        plies.append(places[0])
        finals.append(pack_board(g.board, places[0] & 1, status=status))
    np.savez_compressed(os.path.join(HERE, "greedy_games.npz"), seed=np.uint64(SEED),
                        status=np.array(winners, dtype=np.uint8), plies=np.array(plies, dtype=np.int32),
                        final=np.ascontiguousarray(np.array(finals, dtype=np.uint64).T))
    print("greedy_games: %d games, mean plies %.2f, P1 %d P2 %d stopped %d" % (
        n_games, np.mean(plies), winners.count(1), winners.count(2), winners.count(3)))


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("env", "all"):
        gen_env()
    if what in ("greedy", "all"):
        gen_greedy_games()
    if what in ("mcts", "all"):
        import gen_golden_mcts
        gen_golden_mcts.main()
