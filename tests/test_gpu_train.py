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
