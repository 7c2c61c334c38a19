"""Coordinate helpers with the reference's names and meaning (board_utils.py:3-21); host logic only.

Two coordinate systems describe the 7x7 rhombus.  "np index" (i, j) is the array cell the kernels use
(cell id = 7*i + j).  "Human" (row, col) numbers the 13 anti-diagonals from the bottom tip (row 1) to the top
tip (row 13) and counts cells along a diagonal from 1; row BOARD_WIDTH is the long middle diagonal."""
from .config import BOARD_HEIGHT, BOARD_WIDTH

_MID = BOARD_WIDTH          # human row of the main anti-diagonal


def np_index_to_human_coord(coord):
    """(i, j) -> (row, col): the row is the signed offset from the middle diagonal, the column counts from the
    edge the diagonal starts on, which is whichever index is smaller."""
    i, j = coord
    offset = i - j
    return _MID + offset, 1 + (j if offset > 0 else i)


def human_coord_to_np_index(coord):
    """(row, col) -> (i, j), inverse of np_index_to_human_coord for every on-board cell."""
    row, col = coord
    offset = row - _MID
    along = col - 1
    if offset > 0:
        return along + offset, along
    return along, along - offset


def is_valid_pos(i, j):
    """True when (i, j) lies on the 7x7 array."""
    return (0 <= i < BOARD_HEIGHT) and (0 <= j < BOARD_WIDTH)


def convert_np_to_human_moves(np_moves):
    """{from: [to, ...]} keyed by np index -> the same mapping in human coordinates."""
    out = {}
    for origin, targets in np_moves.items():
        out[np_index_to_human_coord(origin)] = list(map(np_index_to_human_coord, targets))
    return out
