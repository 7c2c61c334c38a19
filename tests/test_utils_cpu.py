"""Host-side helpers that need no GPU: the mirror augmentation (utils.py:77-97 done right), coordinate helpers."""
import numpy as np


def test_augment_mirrors_board_and_pi_consistently():
    from chinesecheckersagent_b200 import utils
    rng = np.random.default_rng(0)
    n = 5
    bx = np.zeros((n, 7, 7, 7)); py = np.zeros((n, 294)); vy = rng.integers(-1, 2, n)
    # put checker id k of the side to move at a random cell and give its move to another random cell all the probability
    for i in range(n):
        k, (r, c), (tr, tc) = int(rng.integers(0, 6)), rng.integers(0, 7, 2), rng.integers(0, 7, 2)
        bx[i, r, c, 0] = k + 1
        py[i, k * 49 + tr * 7 + tc] = 1.0
    ax, ap, av = utils.augment_train_data(bx, py, vy)
    assert ax.shape == (2 * n, 7, 7, 7) and ap.shape == (2 * n, 294) and av.shape == (2 * n,)
    assert np.array_equal(ax[:n], bx) and np.array_equal(ap[:n], py) and np.array_equal(av[n:], vy)
    for i in range(n):
        # the reference's plane transform (utils.py:86-87)
        for j in range(7):
            assert np.array_equal(ax[n + i][:, :, j], np.fliplr(np.rot90(bx[i][:, :, j])))
        k = int(bx[i][..., 0].max()) - 1
        (r, c), = np.argwhere(bx[i][..., 0] > 0)
        (mr, mc), = np.argwhere(ax[n + i][..., 0] > 0)
        assert (mr, mc) == (6 - c, 6 - r)
        t = int(np.argmax(py[i])) % 49
        mt = int(np.argmax(ap[n + i]))
        assert mt // 49 == k and (mt % 49) // 7 == 6 - t % 7 and (mt % 49) % 7 == 6 - t // 7
    # mirroring twice is the identity; mirror_pi=False reproduces the reference's (pi untouched) behaviour
    bx2, ap2, _ = utils.augment_train_data(ax[n:], ap[n:], av[n:])
    assert np.array_equal(bx2[n:], bx) and np.array_equal(ap2[n:], py)
    _, ap3, _ = utils.augment_train_data(bx, py, vy, mirror_pi=False)
    assert np.array_equal(ap3[n:], py)


def test_coordinate_round_trips():
    from chinesecheckersagent_b200 import board_utils
    for i in range(7):
        for j in range(7):
            row, col = board_utils.np_index_to_human_coord((i, j))
            assert row == i - j + 7 and col == min(i, j) + 1
            assert board_utils.human_coord_to_np_index((row, col)) == (i, j)
