"""Self-play bookkeeping fixture: runs the UNMODIFIED selfplay.selfplay() (build container only) with the
tree search replaced by scripted move pickers, and logs what the reference loop did on every ply: the move,
whether it was an MCTS ply, the tree_tau it was searched with, and how the game ended.

The pickers only decide WHICH legal move is played (greedy / random / oscillating); every rule under test —
opening length, repetition discard, progress counter, tau switch, win, useless-move discard, history
truncation, reward — is executed by the reference's own selfplay.py code."""
import contextlib
import io
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

R = refshim.load()
SP = R.selfplay
M = R.MCTS


class StubModel:
    version = 0

    def predict(self, x):
        return np.full(294, 1 / 294.), 0.0


def make_scripted_mcts(policy, rnd):
    class ScriptedMCTS(M.MCTS):
        """search() keeps the reference's contract (pi over the root's edges, a sampled edge) but picks the
        edge with a scripted policy instead of running simulations."""

        def search(self):
            root = self.root
            player = root.currPlayer
            if policy == "greedy":
                cands = R.player.GreedyPlayer(player).decide_move(root.state, training=True)
                s, e = rnd.choice(cands)
                frm, to = R.board_utils.human_coord_to_np_index(s), R.board_utils.human_coord_to_np_index(e)
                edge = [ed for ed in root.edges if ed.fromPos == frm and ed.toPos == to][0]
            elif policy == "random":
                edge = rnd.choice(root.edges)
            else:                                   # "oscillate": undo the mover's previous move when possible
                hm = root.state.hist_moves
                edge = None
                if len(hm) >= 2:
                    pf, pt = hm[-2]
                    back = [ed for ed in root.edges if ed.fromPos == pt and ed.toPos == pf]
                    edge = back[0] if back else None
                if edge is None:
                    edge = rnd.choice(root.edges)
            cid = root.state.checkers_id[player][edge.fromPos]
            root.pi[R.utils.encode_checker_index(cid, edge.toPos)] = 1.0
            return root.pi, edge
    return ScriptedMCTS


def play(policy, seed):
    rnd = random.Random(seed)
    log = []
    orig_random, orig_move, orig_mcts = SP.make_random_move, SP.make_move, SP.MCTS
    random.seed(seed)

    def logged_random(root):
        node = orig_random(root)
        f, t = node.state.hist_moves[-1]
        log.append((orc.cell(*f), orc.cell(*t), 0, 0))
        random.seed(seed + len(log))       # make_random_move reseeds from the OS (selfplay.py:88); keep it reproducible
        return node

    def logged_move(root, model, tree_tau, play_history):
        node = orig_move(root, model, tree_tau, play_history)
        f, t = node.state.hist_moves[-1]
        log.append((orc.cell(*f), orc.cell(*t), 1, int(tree_tau == R.config.DET_TREE_TAU)))
        return node

    SP.make_random_move, SP.make_move, SP.MCTS = logged_random, logged_move, make_scripted_mcts(policy, rnd)
    buf = io.StringIO()
    try:
        with contextlib.redirect_stdout(buf):
            np.random.seed(seed)
            hist, reward = SP.selfplay(StubModel())
    finally:
        SP.make_random_move, SP.make_move, SP.MCTS = orig_random, orig_move, orig_mcts
    text = buf.getvalue()
    if hist is None:
        status = 3 if "Repetition detected" in text else 4
        n_hist = 0
    else:
        status = 1 if reward == 1 else 2
        n_hist = len(hist)
    return log, status, n_hist


def main():
    games = []
    for policy, seeds in (("greedy", range(12)), ("random", range(4)), ("oscillate", range(6))):
        for s in seeds:
            log, status, n_hist = play(policy, 1000 + s)
            games.append((log, status, n_hist))
            print(policy, s, "plies", len(log), "status", status, "history", n_hist)
    L = max(len(g[0]) for g in games)
    moves = np.zeros((len(games), L, 4), dtype=np.uint8)
    for i, (log, _, _) in enumerate(games):
        moves[i, :len(log)] = np.array(log, dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "selfplay_golden.npz"), moves=moves,
                        n_plies=np.array([len(g[0]) for g in games], dtype=np.int32),
                        status=np.array([g[1] for g in games], dtype=np.uint8),
                        n_history=np.array([g[2] for g in games], dtype=np.int32))


if __name__ == "__main__":
    main()
