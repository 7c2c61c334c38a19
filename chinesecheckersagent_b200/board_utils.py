"""Coordinate helpers, same names and meaning as the reference's board_utils.py:3-21 (host logic)."""
from .config import BOARD_HEIGHT, BOARD_WIDTH


def np_index_to_human_coord(coord):
    i, j = coord
    return i - j + BOARD_WIDTH, min(i, j) + 1                       # board_utils.py:3-7


def human_coord_to_np_index(coord):
    row, col = coord
    return col - 1 + max(0, row - BOARD_WIDTH), col - 1 - min(0, row - BOARD_WIDTH)     # board_utils.py:9-13


def is_valid_pos(i, j):
    return 0 <= i < BOARD_HEIGHT and 0 <= j < BOARD_WIDTH


def convert_np_to_human_moves(np_moves):
    return {np_index_to_human_coord(k): [np_index_to_human_coord(t) for t in v] for k, v in np_moves.items()}
