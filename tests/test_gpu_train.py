"""§8f f1 on the GPU: self-play / greedy data -> training step -> the trained weights run in the CUDA inference
kernels (the loop train.evolve closes, train.py:235-318)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu
WEIGHTS = os.path.join(GOLDEN, "good_model_weights.npz")


def test_generate_train_reload_loop(tmp_path):
    from chinesecheckersagent_b200.data_generators import BatchedGreedyGenerator
    from chinesecheckersagent_b200.engine import Engine
    from chinesecheckersagent_b200.model import ResidualCNN
    from chinesecheckersagent_b200.train import TrainableResidualCNN, train
    eng = Engine(0)
    data = BatchedGreedyGenerator(eng, seed=11).generate_mix(200)
    n = data["board_x"].shape[0]
    assert n > 4000
    w = {k: np.asarray(v, dtype=np.float32) for k, v in np.load(WEIGHTS).items()}
    net = TrainableResidualCNN().load_keras_weights(w).cuda()
    hist = train(net, data["board_x"][:4096], data["pi_y"][:4096], data["v_y"][:4096], epochs=2, seed=3)
    assert hist[-1]["loss"] < hist[0]["loss"]
    path = net.save_weights(str(tmp_path / "version0001.npz"))
    # the trained weights in the CUDA inference kernels (BN folded, fp16 tensor-core path and fp32 path) vs torch eval mode
    infer = ResidualCNN(engine=eng).load_weights(path)
    x = data["board_x"][5000:5000 + 512] if n >= 5512 else data["board_x"][:512]
    with torch.no_grad():                               # float64 reference (cuDNN's fp32 convs default to TF32)
        logits, value = net.double()(x)
    p_ref = torch.softmax(logits, dim=1)
    for kernel, bar in (("simt", 2e-5), ("tc", 1e-2)):      # tc: 16-bit operands (precision itself is pinned in tests/test_gpu_net.py)
        infer.set_kernel(kernel)
        p, v = infer.predict_batch(x)
        assert (p - p_ref).abs().max().item() < bar, kernel
        assert (v - value).abs().max().item() < max(bar, 1e-4), kernel
    eng.close()


def test_evolve_one_iteration_end_to_end(tmp_path):
    """train.evolve (train.py:235-317), one bounded iteration on the batched kernels: self-play with the best net ->
    augmented data-for-iter-0.h5 -> training -> version0000-weights.h5 (a Keras save_weights file the loader reads back)
    -> arena against the best net with the 55 % promotion rule."""
    from chinesecheckersagent_b200 import utils
    from chinesecheckersagent_b200.model import ResidualCNN, read_weight_file
    from chinesecheckersagent_b200.train import evolve
    data_dir, w_dir, logs = str(tmp_path / "data"), str(tmp_path / "weights"), []
    cur, best, it = evolve(WEIGHTS, None, 0, WEIGHTS, max_iterations=1, num_self_play=6, eval_games=4, data_dir=data_dir,
                           weights_dir=w_dir, seed=5, epochs=1, log=logs.append, n_slots=32, num_itr=24, max_iters=400)
    assert it == 1 and cur == os.path.join(w_dir, "version0000-weights.h5") and os.path.exists(cur)
    assert best in (WEIGHTS, cur)
    bx, py, vy = utils.load_train_data(os.path.join(data_dir, "data-for-iter-0.h5"))
    n = len(bx)
    assert n >= 2 * 6 and n % 2 == 0 and bx.shape[1:] == (7, 7, 7) and py.shape == (n, 294) and vy.shape == (n,)
    assert np.allclose(py.sum(1), 1.0, atol=1e-5) and set(np.unique(vy)) <= {-1.0, 1.0}
    half = n // 2                                                    # second half = anti-diagonal mirror of the first
    assert np.array_equal(bx[half:], bx[:half, ::-1, ::-1, :].transpose(0, 2, 1, 3)) and np.array_equal(vy[half:], vy[:half])
    w_old, w_new = read_weight_file(WEIGHTS), read_weight_file(cur)
    assert set(w_old) == set(w_new)
    assert any(not np.array_equal(w_old[k], w_new[k]) for k in w_old)   # it trained
    net = ResidualCNN().load_weights(cur)                               # and the result runs in the CUDA kernels
    p, v = net.predict_batch(torch.from_numpy(bx[:8]).cuda())
    assert torch.isfinite(p).all() and torch.isfinite(v).all()
    assert any("self-play" in m for m in logs) and any("best model" in m for m in logs)
