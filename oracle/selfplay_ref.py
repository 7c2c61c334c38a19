"""Restatement of selfplay.selfplay()'s per-ply bookkeeping (selfplay.py:29-80) — TEST INFRASTRUCTURE.

`replay(moves)` pushes a given move sequence through the rules in the reference's order and reports, per
move, what the reference loop would have done: whether the ply is a random opening ply or an MCTS ply,
the tree_tau the MCTS ply was searched with, and the status after the ply (running / won / discarded).
Board mechanics come from the C oracle (oracle/ccx_oracle.c).  Pinned against the real selfplay.py by
tests/golden/selfplay_golden.npz (tests/golden/gen_golden_selfplay.py)."""
import numpy as np

import oracle as orc

RUNNING, WON_P1, WON_P2, REPETITION, MOVE_LIMIT = 0, 1, 2, 3, 4
INITIAL_RANDOM_MOVES, TOTAL_HIST_MOVES, UNIQUE_DEST_LIMIT = 6, 16, 3          # config.py:15-16,38
PROGRESS_MOVE_LIMIT, TOTAL_MOVES_TILL_TAU0, NUM_CHECKERS = 100, 16, 6           # config.py:29,37,8


def replay(moves, start=None):
    """moves: iterable of (from_cell, to_cell) with cell = 8*r + c.  Returns a list of dicts
    (mcts: bool, tau_det: bool, status: int) — one per move actually consumed — and the final state."""
    st = orc.start_states(1) if start is None else np.array(start, dtype=np.uint64).reshape(8, 1)
    player_progresses = [0, 0]                    # selfplay.py:19
    player_turn = 0
    num_useless_moves = 0
    n_history = 0                                 # len(play_history)
    tau_det = False                               # tree_tau == DET_TREE_TAU
    hist = []                                     # board.hist_moves (capped at 16)
    out = []
    for frm, to in moves:
        mcts = len(hist) >= INITIAL_RANDOM_MOVES                                    # selfplay.py:32
        used_det = tau_det
        if mcts:
            n_history += 1                                                           # selfplay.py:128
        st, winner = orc.apply(st, np.array([frm], dtype=np.uint8), np.array([to], dtype=np.uint8))
        hist.append((frm, to))
        if len(hist) > TOTAL_HIST_MOVES:
            hist.pop(0)
        cur = [hist[i] for i in range(len(hist) - 1, -1, -2)]                        # selfplay.py:41
        dests = set(m[1] for m in cur)
        status = RUNNING
        if len(cur) * 2 >= TOTAL_HIST_MOVES and len(dests) <= UNIQUE_DEST_LIMIT:     # selfplay.py:45-47
            status = REPETITION
        else:
            progress = int(orc.info(st)[0, 1 + player_turn])                         # selfplay.py:50
            if progress > player_progresses[player_turn]:
                num_useless_moves = int(num_useless_moves * (NUM_CHECKERS - 1) / NUM_CHECKERS)
                player_progresses[player_turn] = progress
            else:
                num_useless_moves += 1
            player_turn = 1 - player_turn
            if n_history + INITIAL_RANDOM_MOVES > TOTAL_MOVES_TILL_TAU0:             # selfplay.py:62-65
                tau_det = True
            if int(orc.info(st)[0, 0]):                                              # selfplay.py:67-69
                status = int(orc.info(st)[0, 0])
            elif num_useless_moves >= PROGRESS_MOVE_LIMIT:                            # selfplay.py:72-74
                status = MOVE_LIMIT
        out.append(dict(mcts=mcts, tau_det=used_det if mcts else False, status=status))
        if status != RUNNING:
            break
    return out, st
