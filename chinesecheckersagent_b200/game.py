"""Two-seat match driver with the reference `Game` surface (game.py:8-100): `Game(p1_type, p2_type, verbose,
model1, model2, tree_tau).start(enforce_move_limit)` returns the winning player number or None.

The match rules live in `_Referee` (repetition window and ply cap, game.py:70-92) so the same object can be
compared, ply for ply, with what `k_game_advance` (csrc/ccx_selfplay.cu) does for the batched arena."""
import numpy as np

from .board import Board
from .config import DET_TREE_TAU, PROGRESS_MOVE_LIMIT, TOTAL_HIST_MOVES, UNIQUE_DEST_LIMIT
from .player import AiPlayer, GreedyPlayer

_SEAT_KINDS = {"g": "greedy", "a": "ai"}


def _seat(kind, number, model, tree_tau):
    """Player object for one seat; the kind is matched on its first letter like the reference's CLI strings."""
    name = _SEAT_KINDS.get(str(kind)[:1].lower())
    if name == "greedy":
        return GreedyPlayer(player_num=number)
    if name == "ai":
        return AiPlayer(player_num=number, model=model, tree_tau=tree_tau)
    raise ValueError("HumanPlayer (stdin) is out of scope; use 'greedy' or 'ai'")


class _Referee:
    """Stops a match that is going nowhere.  `window` holds the newest TOTAL_HIST_MOVES destinations of both
    sides; once it is full, the mover's own entries (every second one, newest first) must cover more than
    UNIQUE_DEST_LIMIT distinct cells.  `plies` counts non-winning plies for the optional cap."""

    def __init__(self, capped):
        self.capped = bool(capped)
        self.window = []
        self.plies = 0

    def verdict(self, dest):
        """None to play on, else the message the reference prints for the stop reason."""
        self.window = (self.window + [dest])[-TOTAL_HIST_MOVES:]
        if len(self.window) == TOTAL_HIST_MOVES and len(set(self.window[::-2])) <= UNIQUE_DEST_LIMIT:
            return "Repetition detected: stopping game"
        self.plies += 1
        if self.capped and self.plies >= PROGRESS_MOVE_LIMIT:
            return "Game stopped by reaching progress move limit; Game Discarded"
        return None


class Game:
    def __init__(self, p1_type=None, p2_type=None, verbose=True, model1=None, model2=None, tree_tau=DET_TREE_TAU):
        second_model = model2 if model2 is not None else model1
        self.player_one = _seat(p1_type, 1, model1, tree_tau)
        self.player_two = _seat(p2_type, 2, second_model, tree_tau)
        self.cur_player = self.player_one
        self.next_player = self.player_two
        self.board = Board()
        self.verbose = verbose

    def swap_players(self):
        self.cur_player, self.next_player = self.next_player, self.cur_player

    def _ply(self, ply_index):
        """One decision + placement by the seat to move; returns (destination, winner-or-falsy)."""
        mover = self.cur_player
        src, dst = mover.decide_move(self.board, verbose=self.verbose, total_moves=ply_index)
        return dst, self.board.place(mover.player_num, src, dst)

    def start(self, enforce_move_limit=False):
        np.random.seed()
        referee = _Referee(enforce_move_limit)
        winner, ply_index = None, 0
        while winner is None:
            dest, won = self._ply(ply_index)
            ply_index += 1
            if won:
                winner = won
                break
            stop = referee.verdict(dest)
            if stop is not None:
                print(stop)
                break
            self.swap_players()
        if self.verbose:
            self.board.visualise()
        if winner is not None:
            print("Player {} wins!".format(winner))
        return winner
