"""Golden fixture for the greedy data generator (§8f f3): runs the UNMODIFIED reference
data_generators.GreedyDataGenerator.generate_play() + utils.convert_to_train_data in the build container.

    python tests/golden/gen_golden_datagen.py      ->  tests/golden/datagen_golden.npz

Per kind (normal / random_start / randomised) a few games; per record: the packed position with its true side to
move, the reference's pi, and (normal + random_start only — see data_generators.py deviation note) the
convert_to_train_data row (board_x, v_y).  The GPU test replays the same positions through ccx_greedy_candidates /
ccx_cand_to_pi / ccx_encode and checks pi, board_x and the labelling rule bit for bit."""
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, HERE)
import refshim  # noqa: E402
from gen_golden import pack_board  # noqa: E402

R = refshim.load()
dg = __import__("data_generators")
utils = __import__("utils")
dg.STUCK_TIME_LIMIT = 1e9          # no wall-clock truncation while recording fixtures


def main():
    random.seed(20261017)
    np.random.seed(20261017)
    out = {k: [] for k in ("state", "pi", "kind", "game", "board_x", "v_y", "reward", "has_xy")}
    gi = 0
    for kind, kw in enumerate((dict(), dict(random_start=True), dict(randomised=True))):
        gen = dg.GreedyDataGenerator(**kw)
        for _ in range(6):
            # The generator never resets cur_player between games (data_generators.py:71-72 resets the board only), so
            # whoever made the last game's winning move moves first in the next one.  True side to move of record i:
            # the mover at call time (six random-start plies keep the parity), flipped for randomised games, which
            # drop BOARD_HIST_MOVES = 3 records.
            first = (gen.cur_player.player_num - 1) ^ (1 if kind == 2 else 0)
            hist, reward = gen.generate_play()
            # convert_to_train_data assumes player 1 moves first (utils.py:62): only then are its rows comparable
            p1_first = first == 0
            bx, py, vy = utils.convert_to_train_data([(hist, reward)]) if p1_first else (None, None, None)
            for i, (board, pi) in enumerate(hist):
                to_move = (first + i) & 1
                board._ccx_plies = max(len(board.hist_moves), 3 + i if kind == 2 else 0)
                out["state"].append(pack_board(board, to_move))
                out["pi"].append(np.asarray(pi, dtype=np.float64))
                out["kind"].append(kind); out["game"].append(gi); out["reward"].append(reward)
                if p1_first:
                    out["board_x"].append(np.asarray(bx[i], dtype=np.uint8)); out["v_y"].append(int(vy[i])); out["has_xy"].append(1)
                else:
                    out["board_x"].append(np.zeros((7, 7, 7), np.uint8)); out["v_y"].append(0); out["has_xy"].append(0)
            gi += 1
    np.savez_compressed(os.path.join(HERE, "datagen_golden.npz"),
                        state=np.stack(out["state"], axis=1), pi=np.stack(out["pi"]), kind=np.array(out["kind"], np.int8),
                        game=np.array(out["game"], np.int32), board_x=np.stack(out["board_x"]), v_y=np.array(out["v_y"], np.int8),
                        reward=np.array(out["reward"], np.int8), has_xy=np.array(out["has_xy"], np.int8))
    print("records", len(out["kind"]), "games", gi)


if __name__ == "__main__":
    main()
