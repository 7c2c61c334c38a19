"""Mirror of the reference's game.py `Game` (game.py:8-100) for greedy / AI players."""
from collections import deque

import numpy as np

from .board import Board
from .config import DET_TREE_TAU, PROGRESS_MOVE_LIMIT, TOTAL_HIST_MOVES, UNIQUE_DEST_LIMIT
from .player import AiPlayer, GreedyPlayer


class Game:
    def __init__(self, p1_type=None, p2_type=None, verbose=True, model1=None, model2=None, tree_tau=DET_TREE_TAU):
        def make(kind, num, model):
            k = kind[0].lower()
            if k == 'g':
                return GreedyPlayer(player_num=num)
            if k == 'a':
                return AiPlayer(player_num=num, model=model, tree_tau=tree_tau)
            raise ValueError("HumanPlayer (stdin) is out of scope; use 'greedy' or 'ai'")
        self.player_one = make(p1_type, 1, model1)
        self.player_two = make(p2_type, 2, model1 if model2 is None else model2)
        self.cur_player, self.next_player = self.player_one, self.player_two
        self.verbose = verbose
        self.board = Board()

    def swap_players(self):
        self.cur_player, self.next_player = self.next_player, self.cur_player

    def start(self, enforce_move_limit=False):
        np.random.seed()
        total_moves = 0
        history_dests = deque()
        num_moves = 0
        while True:
            move_from, move_to = self.cur_player.decide_move(self.board, verbose=self.verbose, total_moves=total_moves)
            winner = self.board.place(self.cur_player.player_num, move_from, move_to)      # game.py:65
            total_moves += 1
            if winner:
                break
            if len(history_dests) == TOTAL_HIST_MOVES:
                history_dests.popleft()
            history_dests.append(move_to)
            mine = set(history_dests[i] for i in range(len(history_dests) - 1, -1, -2))     # game.py:78
            if len(history_dests) == TOTAL_HIST_MOVES and len(mine) <= UNIQUE_DEST_LIMIT:
                print('Repetition detected: stopping game')
                winner = None
                break
            num_moves += 1
            if enforce_move_limit and num_moves >= PROGRESS_MOVE_LIMIT:
                print('Game stopped by reaching progress move limit; Game Discarded')
                winner = None
                break
            self.swap_players()
        if self.verbose:
            self.board.visualise()
        if winner is not None:
            print('Player {} wins!'.format(winner))
        return winner
