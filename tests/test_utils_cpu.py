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


def test_combine_train_data_and_label_counts(tmp_path):
    """f4 tools (combine_data.py, count_labels.py, train.py:321-352): pool current examples with saved iterations,
    skip missing files and negative versions, count labels."""
    import numpy as np
    from chinesecheckersagent_b200 import utils
    from chinesecheckersagent_b200.train import combine_prev_iters_train_data
    rng = np.random.default_rng(3)
    def make(n, label):
        return (rng.integers(0, 3, (n, 7, 7, 7)).astype(np.uint8), rng.random((n, 294)).astype(np.float32),
                np.full((n,), label, np.float32))
    d = str(tmp_path)
    sets = {v: make(3 + v, (-1.0, 1.0)[v & 1]) for v in (0, 1, 3)}
    for v, (bx, py, vy) in sets.items():
        utils.save_train_data(bx, py, vy, version=v, directory=d)
    cur = make(2, 1.0)
    bx, py, vy, used = utils.combine_train_data(*cur, -2, 3, d)                     # versions -2, -1 ignored; 2 missing
    assert used == 4 and bx.shape == (2 + 3 + 4 + 6, 7, 7, 7) and py.shape == (15, 294) and vy.shape == (15,)
    assert np.array_equal(bx[:2], cur[0]) and np.array_equal(bx[2:5], sets[0][0]) and np.array_equal(py[9:], sets[3][1])
    assert utils.count_items(vy) == {-1.0: 3, 1.0: 12}
    assert utils.get_train_label_count("%s/data-for-iter-1.h5" % d) == {1.0: 4}
    assert utils.combine_train_data([], [], [], 5, 6, d) == ([], [], [], 0)
    bx2, _, vy2, used2 = combine_prev_iters_train_data(*cur, 1, save_dir=d)        # PAST_ITER_COUNT = 1 -> file 0 only
    assert used2 == 2 and len(bx2) == 5 and utils.count_items(vy2) == {-1.0: 3, 1.0: 2}
    _, _, _, used3 = combine_prev_iters_train_data([], [], [], 0, save_dir=d)       # nothing before iteration 0
    assert used3 == 0
