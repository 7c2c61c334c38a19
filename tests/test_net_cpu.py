"""CPU-only checks around the net: the HDF5-free weight fixture, BN folding / packing (host logic) and
the NumPy restatement's known answers.  PARITY UNPINNED vs Keras (see oracle/net_ref.py)."""
import os

import numpy as np

import net_ref
import oracle as orc
from conftest import GOLDEN


def weights():
    return dict(np.load(os.path.join(GOLDEN, "good_model_weights.npz")))


def test_weight_fixture_shape_inventory():
    w = weights()
    assert len(w) == 186 and sum(v.size for v in w.values()) == 249852
    assert w["conv2d_1/kernel"].shape == (3, 3, 7, 64) and w["policy_head/kernel"].shape == (400, 294)


def test_start_position_known_answer():
    p, v = net_ref.predict(weights(), orc.encode(orc.start_states(1)), np.float32)
    assert abs(float(v[0]) - (-0.043674)) < 1e-5
    assert list(np.argsort(-p[0])[:5]) == [76, 117, 101, 60, 143]


def test_golden_outputs_reproduce():
    g = np.load(os.path.join(GOLDEN, "net_golden.npz"))
    logits, v = net_ref.forward(weights(), g["planes"], np.float64)
    assert np.allclose(logits, g["logits"], atol=1e-9) and np.allclose(v, g["v"], atol=1e-12)


def test_folded_packed_weights_equal_unfolded_graph():
    """pack_weights (BN folded into conv) evaluated with plain matmuls == net_ref's unfolded graph."""
    from chinesecheckersagent_b200.model import pack_weights
    w = weights()
    blob = pack_weights(w).astype(np.float64)
    assert blob.size == 244920
    g = np.load(os.path.join(GOLDEN, "net_golden.npz"))
    x = g["planes"][:32].astype(np.float64)
    off = [0]

    def take(n):
        a = blob[off[0]:off[0] + n]; off[0] += n
        return a
    relu = lambda t: np.maximum(t, 0)
    # conv1 (valid)
    W, b = take(63 * 64).reshape(63, 64), take(64)
    cols = np.concatenate([x[:, dy:dy + 5, dx:dx + 5, :] for dy in range(3) for dx in range(3)], axis=-1)
    a = relu(cols @ W + b)
    for _ in range(9):
        Wa, ba = take(64 * 32).reshape(64, 32), take(32)
        Wb, bb = take(288 * 32).reshape(288, 32), take(32)
        Wc, bc = take(32 * 64).reshape(32, 64), take(64)
        m = relu(a @ Wa + ba)
        mp = np.pad(m, ((0, 0), (1, 1), (1, 1), (0, 0)))
        cols = np.concatenate([mp[:, dy:dy + 5, dx:dx + 5, :] for dy in range(3) for dx in range(3)], axis=-1)
        m = relu(cols @ Wb + bb)
        a = relu(m @ Wc + bc + a)
    Wp, bp = take(64 * 16).reshape(64, 16), take(16)
    Wd, bd = take(400 * 294).reshape(400, 294), take(294)
    logits = relu(a @ Wp + bp).reshape(len(x), -1) @ Wd + bd
    Wv, bv = take(64).reshape(64, 1), take(1)
    W1, b1 = take(25 * 32).reshape(25, 32), take(32)
    Wh, bh = take(32).reshape(32, 1), take(1)
    v = np.tanh(relu(relu(a @ Wv + bv).reshape(len(x), -1) @ W1 + b1) @ Wh + bh)[:, 0]
    assert off[0] == blob.size
    assert np.allclose(logits, g["logits"][:32], atol=2e-4) and np.allclose(v, g["v"][:32], atol=2e-5)


def test_board_row_major_tile_rows_make_every_3x3_tap_a_row_shift():
    """host model of k_net_trunk_accm's ROWMAJ operand buffer (csrc/ccx_net_tc.cu, namespace acm): rows r = y*24 + p*6 + x of a
    four-position tile behind 25 zero guard rows, cells x = 5 and rows >= 120 never written.  For every output cell and every tap
    (dy, dx) of the 3x3 'same' conv (model.py:128-132) the row r + 24*dy + dx must hold the input cell (y+dy, x+dx) of the SAME
    position, or a zero when that cell is off the board — which is what lets one buffer serve all nine taps."""
    G, ROWS = 25, 128
    rng = np.random.default_rng(5)
    act = rng.standard_normal((4, 5, 5))                       # one channel is enough: the layout is per row
    buf = np.zeros(G + ROWS + G)
    for p in range(4):
        for y in range(5):
            for x in range(5):
                buf[G + y * 24 + p * 6 + x] = act[p, y, x]
    for p in range(4):
        for y in range(5):
            for x in range(5):
                r = y * 24 + p * 6 + x
                for dy in (-1, 0, 1):
                    for dx in (-1, 0, 1):
                        src = G + r + 24 * dy + dx
                        assert 0 <= src < len(buf)
                        yy, xx = y + dy, x + dx
                        want = act[p, yy, xx] if 0 <= yy < 5 and 0 <= xx < 5 else 0.0
                        assert buf[src] == want, (p, y, x, dy, dx)
    # the MMA reads 128 consecutive rows from start row G + 24*dy + dx: always inside the buffer
    assert G - 24 - 1 >= 0 and G + 24 + 1 + ROWS <= len(buf)
