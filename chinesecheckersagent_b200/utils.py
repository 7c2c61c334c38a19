"""Mirror of the hot-path part of the reference's utils.py: to_model_input (utils.py:101-160, on the GPU),
encode/decode_checker_index (utils.py:164-183), softmax (utils.py:187-192), get_p1_winloss_reward
(utils.py:34-44) and convert_to_train_data (utils.py:60-73)."""
import numpy as np

from . import engine as _engine
from .config import BOARD_HEIGHT, BOARD_WIDTH, DTYPE_U8, PLAYER_ONE, PLAYER_TWO, REWARD


def to_model_input(board, cur_player):
    """(7,7,7) float64 channels-last, like the reference; computed by the fused encoder kernel."""
    return board._host().encode(board._pack(cur_player - 1)).astype(np.float64)


def encode_checker_index(checker_id, coord):
    return checker_id * BOARD_WIDTH * BOARD_HEIGHT + coord[0] * BOARD_WIDTH + coord[1]


def decode_checker_index(model_output_index):
    checker_id = model_output_index // (BOARD_WIDTH * BOARD_HEIGHT)
    offset = model_output_index % (BOARD_WIDTH * BOARD_HEIGHT)
    return checker_id, (offset // BOARD_WIDTH, offset % BOARD_WIDTH)


def softmax(input):
    input = np.copy(input).astype('float64')
    input -= np.max(input, axis=-1, keepdims=True)
    exps = np.exp(input)
    return exps / np.sum(exps, axis=-1, keepdims=True)


def get_p1_winloss_reward(board, winner=None):
    winner = winner or board.check_win()
    if winner == PLAYER_ONE:
        return REWARD['win']
    if winner == PLAYER_TWO:
        return REWARD['lose']
    return REWARD['draw']


def convert_to_train_data(self_play_games):
    board_x, pi_y, v_y = [], [], []
    for history, reward in self_play_games:
        curr_player = PLAYER_ONE
        for board, pi in history:
            board_x.append(to_model_input(board, curr_player))
            pi_y.append(pi)
            v_y.append(reward)
            reward = -reward
            curr_player = PLAYER_ONE + PLAYER_TWO - curr_player
    return board_x, pi_y, v_y


def save_train_data(board_x, pi_y, v_y, version, directory="generated-training-data/", prefix="data-for-iter-"):
    """utils.py:48-56 — `<dir>/data-for-iter-<version>.h5` with datasets board_x (N,7,7,7), pi_y (N,294), v_y (N,)
    (written with the package's own HDF5 writer; no h5py needed)."""
    import os

    from . import h5lite
    os.makedirs(directory, exist_ok=True)
    path = os.path.join(directory, "%s%s.h5" % (prefix, version))
    to_np = lambda t: t.detach().cpu().numpy() if hasattr(t, "detach") else np.asarray(t)
    return h5lite.write_tree(path, {"board_x": to_np(board_x), "pi_y": to_np(pi_y), "v_y": to_np(v_y)})


def load_train_data(path):
    """(board_x, pi_y, v_y) from a data-for-iter file (train.py:321-352 reads them back with h5py)."""
    from . import h5lite
    t = h5lite.read_tree(path)
    return t["/board_x"], t["/pi_y"], t["/v_y"]


def augment_train_data(board_x, pi_y, v_y, mirror_pi=True):
    """utils.augment_train_data (utils.py:77-97): doubles the data with the board mirrored along the anti-diagonal
    (np.fliplr(np.rot90(plane)) = cell (r, c) -> (6 - c, 6 - r), the game's mirror symmetry).  The reference mirrors the
    boards but leaves pi untouched, which mislabels every mirrored example (SURVEY 8a quirk ii); here pi is mirrored with the
    board (policy index id*49 + r*7 + c -> id*49 + (6-c)*7 + (6-r)) unless mirror_pi=False asks for the reference's behaviour.
    Accepts lists / arrays (N,7,7,7), (N,294), (N,); returns arrays of twice the length."""
    bx, py, vy = np.asarray(board_x), np.asarray(pi_y), np.asarray(v_y)
    mb = bx[:, ::-1, ::-1, :].transpose(0, 2, 1, 3)               # out[r', c'] = in[6 - c', 6 - r']
    mp = py
    if mirror_pi:
        p = py.reshape(len(py), 6, BOARD_HEIGHT, BOARD_WIDTH)
        mp = p[:, :, ::-1, ::-1].transpose(0, 1, 3, 2).reshape(len(py), -1)
    return np.concatenate([bx, mb]), np.concatenate([py, mp]), np.concatenate([vy, vy])


# ---- data merge / label-count tools (combine_data.py, count_labels.py, train.py:321-352) ----
def combine_train_data(board_x, pi_y, v_y, first_version, last_version, save_dir="generated-training-data",
                       pref="data-for-iter-"):
    """Pool the given examples (may be empty) with the saved files `<save_dir>/<pref><i>.h5` for every
    i in [first_version, last_version] that is >= 0 and exists (missing files are reported and skipped, as
    combine_data.py:13-25 does).  Returns (board_x, pi_y, v_y, number of sources pooled); ([], [], [], 0) when
    there is nothing at all."""
    import os
    parts = []
    if len(board_x) and len(pi_y) and len(v_y):
        parts.append((np.asarray(board_x), np.asarray(pi_y), np.asarray(v_y)))
    for version in range(max(int(first_version), 0), int(last_version) + 1):
        path = "%s/%s%d.h5" % (save_dir, pref, version)
        if os.path.exists(path):
            parts.append(tuple(np.asarray(a) for a in load_train_data(path)))
        else:
            print("{} does not exist!".format(path))
    if not parts:
        return [], [], [], 0
    boards, pis, vs = zip(*parts)
    return np.concatenate(boards, axis=0), np.concatenate(pis, axis=0), np.concatenate([np.ravel(v) for v in vs]), len(parts)


def count_items(v_y):
    """count_labels.count_items: {label value: occurrences} of a v_y vector."""
    values, counts = np.unique(np.asarray(v_y), return_counts=True)
    return {v.item(): int(c) for v, c in zip(values, counts)}


def get_train_label_count(path):
    """count_labels.get_train_label_count: label histogram of one data-for-iter file."""
    return count_items(load_train_data(path)[2])


# ---- small host helpers the reference's scripts call (utils.py:12-31) ------------------------------------------------
def stress_message(message, extra_newline=False):
    """The message framed by two '=' rules of its own length (greedy_vs_greedy.py:16, ai_vs_ai.py:36 print game banners with it)."""
    rule = '=' * len(message)
    pad = '\n' if extra_newline else ''
    print(pad + rule + '\n' + message + '\n' + rule + pad)


def find_version_given_filename(filename, prefixes=("version", "greedy-model")):
    """4-digit version inside `<prefix>NNNN[-weights].h5` (ai_vs_ai.py:21 tags model.version with it), -1 when absent."""
    import re
    m = re.search(r'(?:%s)(\d{4})(?:-weights)?\.(?:h5|npz)' % '|'.join(re.escape(p) for p in prefixes), str(filename))
    if m is None:
        print('No 4-digit version number found in filename "{}"!'.format(filename))
        return -1
    return int(m.group(1))


def cur_time():
    import datetime
    return datetime.datetime.now().strftime("%Y-%m-%d %H:%M:%S")
