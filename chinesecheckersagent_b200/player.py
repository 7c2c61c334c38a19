"""Mirror of the reference's player.py: GreedyPlayer (player.py:67-129, both the deterministic-heuristic and the stochastic branch)
and AiPlayer (player.py:133-166).
HumanPlayer (interactive stdin) is out of scope."""
import random

import numpy as np

from . import board_utils
from . import engine as _engine
from .config import DET_TREE_TAU, PLAYER_ONE, TOTAL_MOVES_TILL_TAU0
from .MCTS import MCTS, Node


class GreedyPlayer:
    def __init__(self, player_num, stochastic=False):
        self.player_num, self.stochastic = player_num, stochastic

    def _decide_stochastic(self, board):
        """player.py:77-97: forward moves sampled in proportion to the rows they advance, a uniform backward / sideways
        move when nothing advances.  Move lists come from the device movegen; the sampling is host logic like the reference's."""
        human = board_utils.convert_np_to_human_moves(board.get_valid_moves(self.player_num))
        prior, forward, backward = [], [], []
        for start in human:
            for end in human[start]:
                dist = end[0] - start[0]
                if self.player_num == PLAYER_ONE:
                    dist = -dist
                if dist > 0:
                    forward.append((start, end)); prior.append(dist)
                else:
                    backward.append((start, end))
        if not forward:
            return random.choice(backward)
        p = np.array(prior) / sum(prior)
        return forward[np.random.choice(len(forward), p=p)]

    def decide_move(self, board, verbose=False, training=False, total_moves=None):
        if self.stochastic:
            pick_start, pick_end = self._decide_stochastic(board)
            if verbose:
                board.visualise(cur_player=self.player_num)
                print('GreedyPlayer moved from {} to {}\n'.format(pick_start, pick_end))
            return board_utils.human_coord_to_np_index(pick_start), board_utils.human_coord_to_np_index(pick_end)
        masks = board._host().greedy(board._pack(self.player_num - 1))
        moves = []
        for cid in range(6):
            m = int(masks[cid])
            for b in range(55):
                if (m >> b) & 1:
                    moves.append((board.checkers_pos[self.player_num][cid], (b >> 3, b & 7)))
        if not moves:
            raise ValueError("max() arg is an empty sequence")              # what player.py:113 raises
        if training:                                                        # player.py:117-118 (human coordinates)
            return [(board_utils.np_index_to_human_coord(s), board_utils.np_index_to_human_coord(e)) for s, e in moves]
        pick_start, pick_end = random.choice(moves)                         # player.py:121
        if verbose:
            board.visualise(cur_player=self.player_num)
            print('GreedyPlayer moved from {} to {}\n'.format(board_utils.np_index_to_human_coord(pick_start),
                                                              board_utils.np_index_to_human_coord(pick_end)))
        return pick_start, pick_end


class AiPlayer:
    def __init__(self, player_num, model, tree_tau):
        self.player_num, self.model, self.tree_tau = player_num, model, tree_tau

    def decide_move(self, board, verbose=False, total_moves=None):
        if verbose:
            board.visualise(cur_player=self.player_num)
            print('Facing the board above, Ai Version {} is thinking.'.format(self.model.version))
        node = Node(board, self.player_num)
        if total_moves is not None and total_moves > TOTAL_MOVES_TILL_TAU0:      # player.py:152-155
            self.tree_tau = DET_TREE_TAU
        tree = MCTS(node, self.model, tree_tau=self.tree_tau)
        pi, sampled_edge = tree.search()                                         # player.py:157-158
        return sampled_edge.fromPos, sampled_edge.toPos
