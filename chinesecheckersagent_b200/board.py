"""Drop-in mirror of the reference's `Board` (board.py:9-288): same constructor, methods, attributes and
return types, with every rule evaluated by libccx.so's CUDA kernels on a batch of one.

Differences a caller can observe (documented in INTEGRATION.md): destination lists come back in canonical
order (ascending r*7+c) instead of the reference's walk-then-DFS order, and `.board` is a read-only
snapshot rebuilt from the packed state.  There is no CPU path: constructing a Board needs a CUDA device."""
import copy
import ctypes
from collections import deque

import numpy as np
import torch

from . import engine as _engine
from .config import (BOARD_HEIGHT, BOARD_HIST_MOVES, BOARD_WIDTH, NUM_CHECKERS, PLAYER_ONE, PLAYER_TWO, STATE_WORDS,
                     TOTAL_HIST_MOVES)

_default_engines = {}


def default_engine():
    """the engine of torch's CURRENT CUDA device (one per device: rank r of a multi-GPU job gets device r's)"""
    dev = torch.cuda.current_device() if torch.cuda.is_available() else 0
    if dev not in _default_engines:
        _default_engines[dev] = _engine.Engine(dev)
    return _default_engines[dev]


class _HostCalls:
    """One-board calls through the host-buffer C-ABI (`ccx_*_host`: H2D, kernel, D2H and the stream sync inside ONE C call, no
    torch tensors): what keeps a reference script that was only re-pointed at this package usable — ~30 us per Board method
    instead of a tensor allocation + three torch round trips.  Buffers are per engine and reused."""
    _by_engine = {}

    def __init__(self, eng):
        self.eng = eng
        self.masks = np.zeros((6, 1), dtype=np.uint64)
        self.info = np.zeros((1, 5), dtype=np.int16)
        self.frm = np.zeros(1, dtype=np.uint8)
        self.to = np.zeros(1, dtype=np.uint8)
        self.winner = np.zeros(1, dtype=np.uint8)
        self.planes = np.zeros((1, 7, 7, 7), dtype=np.uint8)

    @classmethod
    def of(cls, eng):
        hc = cls._by_engine.get(id(eng))
        if hc is None or hc.eng is not eng:
            hc = cls._by_engine[id(eng)] = cls(eng)
        return hc

    @staticmethod
    def _ptr(a):
        return ctypes.c_void_p(a.ctypes.data)

    def movegen(self, st):
        self.eng.call("ccx_movegen_host", 1, self._ptr(st), self._ptr(self.masks))
        return self.masks[:, 0]

    def greedy(self, st):
        """filtered_best_moves masks (player.py:99-118)"""
        self.eng.call("ccx_greedy_candidates_host", 1, self._ptr(st), self._ptr(self.masks))
        return self.masks[:, 0]

    def query(self, st):
        self.eng.call("ccx_info_host", 1, self._ptr(st), self._ptr(self.info))
        return self.info[0]

    def apply(self, st, frm, to):
        self.frm[0], self.to[0] = frm, to
        self.eng.call("ccx_apply_host", 1, self._ptr(st), self._ptr(self.frm), self._ptr(self.to), self._ptr(self.winner))
        return int(self.winner[0])

    def encode(self, st):
        self.eng.call("ccx_encode_host", 1, self._ptr(st), self._ptr(self.planes), 0)
        return self.planes[0]


def _cell(pos):
    return 8 * int(pos[0]) + int(pos[1])


def _rc(cell):
    return int(cell) >> 3, int(cell) & 7


START_P1 = [(6, 0), (5, 0), (6, 1), (4, 0), (5, 1), (6, 2)]      # board.py:43-44
START_P2 = [(0, 6), (1, 6), (0, 5), (2, 6), (1, 5), (0, 4)]      # board.py:45-46


class Board:
    def __init__(self, randomised=False, engine=None):
        self._eng = engine or default_engine()
        self.directions = [(-1, 0), (0, 1), (1, 1), (1, 0), (0, -1), (-1, -1)]     # board.py:33-40
        self.checkers_pos = [None, dict(enumerate(START_P1)), dict(enumerate(START_P2))]
        self.hist_moves = deque()
        self._plies = 0
        if randomised:
            self.randomise_initial_state()
        self._sync_ids()

    @classmethod
    def from_packed(cls, words, engine=None):
        """Board from uint64 state words (include/ccx.h layout; words 0-4 suffice).  hist_moves holds the last two moves —
        all that utils.to_model_input reads (utils.py:135-150)."""
        b = cls.__new__(cls)
        b._eng = engine or default_engine()
        b.directions = [(-1, 0), (0, 1), (1, 1), (1, 0), (0, -1), (-1, -1)]
        c1, c2, meta = int(words[2]), int(words[3]), int(words[4])
        b.checkers_pos = [None, {i: _rc((c1 >> (8 * i)) & 0xFF) for i in range(NUM_CHECKERS)},
                          {i: _rc((c2 >> (8 * i)) & 0xFF) for i in range(NUM_CHECKERS)}]
        b._plies = (meta >> 32) & 0xFFFF
        b.hist_moves = deque()
        for k in (1, 0):
            f, t = (meta >> (16 * k)) & 0xFF, (meta >> (16 * k + 8)) & 0xFF
            if f != 0xFF and k < b._plies:
                b.hist_moves.append((_rc(f), _rc(t)))
        b._sync_ids()
        return b

    # -- packed state <-> python attributes ------------------------------------------------------------
    def _sync_ids(self):
        self.checkers_id = [None, {p: i for i, p in self.checkers_pos[1].items()},
                            {p: i for i, p in self.checkers_pos[2].items()}]

    def _pack(self, to_move):
        w = np.zeros((STATE_WORDS, 1), dtype=np.uint64)
        for pl in (1, 2):
            occ = cells = 0
            for i in range(NUM_CHECKERS):
                c = _cell(self.checkers_pos[pl][i])
                occ |= 1 << c
                cells |= c << (8 * i)
            w[pl - 1, 0], w[pl + 1, 0] = occ, cells
        hm = list(self.hist_moves)
        meta = 0
        for k in range(2):
            if k < len(hm):
                f, t = hm[-1 - k]
                meta |= _cell(f) << (16 * k) | _cell(t) << (16 * k + 8)
            else:
                meta |= 0xFFFF << (16 * k)
        meta |= (self._plies & 0xFFFF) << 32 | (to_move & 1) << 48
        hist = [(1 << 64) - 1] * 2
        for k in range(len(hm)):
            hist[k >> 3] = (hist[k >> 3] & ~(0xFF << (8 * (k & 7)))) | (_cell(hm[-1 - k][1]) << (8 * (k & 7)))
        w[4, 0], w[5, 0], w[6, 0] = meta, hist[0], hist[1]
        return w

    def _env(self, player):
        return _engine.BatchedEnv(1, engine=self._eng, state=self._pack(player - 1))

    def _host(self):
        return _HostCalls.of(self._eng)

    def packed_state(self, cur_player):
        """uint64[8] in the include/ccx.h layout with `cur_player` to move."""
        return self._pack(cur_player - 1)[:, 0]

    @property
    def board(self):
        """(7,7,3) uint8: plane 0 current position, planes 1-2 the two previous ones (board.py:19-26, 243)."""
        b = np.zeros((BOARD_WIDTH, BOARD_HEIGHT, BOARD_HIST_MOVES), dtype="uint8")
        for pl in (1, 2):
            for p in self.checkers_pos[pl].values():
                b[p[0], p[1], 0] = pl
        hm = list(self.hist_moves)
        cur = b[:, :, 0].copy()
        for k in range(1, BOARD_HIST_MOVES):
            if k > self._plies or k > len(hm):
                break
            f, t = hm[-k]
            cur[f], cur[t] = cur[t], cur[f]
            b[:, :, k] = cur
        return b

    # -- board.py API --------------------------------------------------------------------------------------
    def randomise_initial_state(self):
        """board.py:61-85 — host RNG (np.random) like the reference, so np.random.seed() keeps its meaning."""
        chosen = np.random.choice(BOARD_WIDTH * BOARD_HEIGHT, size=NUM_CHECKERS * 2, replace=False)
        pos = [(int(i) // BOARD_WIDTH, int(i) % BOARD_WIDTH) for i in chosen]
        self.checkers_pos = [None, dict(enumerate(pos[:NUM_CHECKERS])), dict(enumerate(pos[NUM_CHECKERS:]))]
        self._sync_ids()

    def check_win(self):
        return int(self._host().query(self._pack(0))[0])                          # board.py:89-111

    def player_progress(self, player_id):
        return int(self._host().query(self._pack(0))[player_id])                  # board.py:254-266

    def player_forward_distance(self, player_id):
        return int(self._host().query(self._pack(0))[2 + player_id])              # board.py:270-288

    def get_valid_moves(self, cur_player):
        """board.py:215-222: {checker position: [destinations]} keyed in checker-id order."""
        masks = self._host().movegen(self._pack(cur_player - 1))
        out = {}
        for i in range(NUM_CHECKERS):
            m = int(masks[i])
            out[self.checkers_pos[cur_player][i]] = [_rc(b) for b in range(55) if (m >> b) & 1]
        return out

    def valid_checker_moves(self, cur_player, checker_pos):
        return self.get_valid_moves(cur_player)[tuple(checker_pos)]               # board.py:139-162

    def place(self, cur_player, origin_pos, dest_pos):
        """board.py:226-250: moves the checker, shifts history, returns check_win()."""
        origin_pos, dest_pos = tuple(int(x) for x in origin_pos), tuple(int(x) for x in dest_pos)
        st = self._pack(cur_player - 1)
        winner = self._host().apply(st, _cell(origin_pos), _cell(dest_pos))
        st = st[:, 0]
        for pl in (1, 2):
            for i in range(NUM_CHECKERS):
                self.checkers_pos[pl][i] = _rc((int(st[pl + 1]) >> (8 * i)) & 0xFF)
        self._sync_ids()
        if len(self.hist_moves) == TOTAL_HIST_MOVES:
            self.hist_moves.popleft()
        self.hist_moves.append((origin_pos, dest_pos))
        self._plies += 1
        return winner

    def visualise(self, cur_player=None, gap_btw_checkers=3):
        """board.py:115-135"""
        print('=' * 75)
        print('Current Status:' + ' ' * 40 + 'Current Player: {}\n'.format(cur_player))
        cur_board = self.board[:, :, 0]
        visual_width = BOARD_WIDTH * (gap_btw_checkers + 1) - gap_btw_checkers
        visual_height = BOARD_HEIGHT * 2 - 1
        leading_spaces = visual_width // 2
        for i in range(1, visual_height + 1):
            num_slots = i if i <= BOARD_WIDTH else visual_height - i + 1
            print('\tRow {:2}{}'.format(i, ' ' * 8), end='')
            print(' ' * (leading_spaces - (num_slots - 1) * ((gap_btw_checkers + 1) // 2)), end='')
            print((' ' * gap_btw_checkers).join(map(str, cur_board.diagonal(BOARD_WIDTH - i))), end='\n\n')
        print('=' * 75)

    def __deepcopy__(self, memo):
        b = Board.__new__(Board)
        b._eng = self._eng                          # the engine handle is shared, never copied
        b.directions = list(self.directions)
        b.checkers_pos = [None, dict(self.checkers_pos[1]), dict(self.checkers_pos[2])]
        b.hist_moves = deque(self.hist_moves)
        b._plies = self._plies
        b._sync_ids()
        return b
