"""§8f f1 on CPU: the trainable PyTorch net is the same function as the restated Keras graph (oracle/net_ref.py),
weights round-trip through the Keras tensor layout, and the loss / optimiser follow model.py:58-87."""
import os

import numpy as np
import torch

import net_ref
from conftest import GOLDEN

WEIGHTS = os.path.join(GOLDEN, "good_model_weights.npz")


def load():
    from chinesecheckersagent_b200.train import TrainableResidualCNN
    w = {k: np.asarray(v, dtype=np.float32) for k, v in np.load(WEIGHTS).items()}
    return TrainableResidualCNN().double().load_keras_weights({k: v.astype(np.float64) for k, v in w.items()}), w


def test_forward_equals_restatement_and_weights_round_trip():
    model, w = load()
    gold = np.load(os.path.join(GOLDEN, "net_golden.npz"))
    planes = gold["planes"][:64]
    model.eval()
    with torch.no_grad():
        logits, value = model(torch.from_numpy(planes))
    ref_l, ref_v = net_ref.forward(w, planes, np.float64)
    assert np.abs(logits.numpy() - ref_l).max() < 1e-9 and np.abs(value.numpy() - ref_v).max() < 1e-9
    back = model.float().keras_weights()
    assert set(back) == set(w)
    for k in w:
        assert back[k].shape == w[k].shape and np.array_equal(back[k], w[k]), k


def test_loss_terms_and_l2():
    from chinesecheckersagent_b200.config import REG_CONST
    from chinesecheckersagent_b200.train import loss_terms
    model, w = load()
    model.eval()
    gold = np.load(os.path.join(GOLDEN, "net_golden.npz"))
    planes = torch.from_numpy(gold["planes"][:32])
    pi = torch.zeros(32, 294, dtype=torch.float64); pi[:, 5] = 0.25; pi[:, 70] = 0.75
    v = torch.ones(32, dtype=torch.float64)
    total, ce, mse, l2 = loss_terms(model, planes, pi, v)
    ref_l, ref_v = net_ref.forward(w, gold["planes"][:32], np.float64)
    logp = np.log(net_ref.softmax64(ref_l))
    assert abs(float(ce.detach()) - float(-(pi.numpy() * logp).sum(1).mean())) < 1e-9
    assert abs(float(mse.detach()) - float(((ref_v - 1.0) ** 2).mean())) < 1e-9
    kern = sum(float((w[k].astype(np.float64) ** 2).sum()) for k in w if k.endswith("/kernel"))
    assert abs(float(l2.detach()) - REG_CONST * kern) < 1e-6 * REG_CONST * kern
    assert abs(float(total.detach()) - float((ce + mse + l2).detach())) < 1e-9


def test_nesterov_update_is_keras_formula():
    """Keras SGD(nesterov): v = m v - lr g; w = w + m v - lr g  (keras/optimizers.py) on a toy parameter"""
    p = torch.nn.Parameter(torch.tensor([1.0, -2.0], dtype=torch.float64))
    opt = torch.optim.SGD([p], lr=1e-4, momentum=0.9, nesterov=True)
    wk, vk = np.array([1.0, -2.0]), np.zeros(2)
    for _ in range(5):
        opt.zero_grad()
        (p ** 3).sum().backward()
        g = 3 * wk ** 2
        opt.step()
        vk = 0.9 * vk - 1e-4 * g
        wk = wk + 0.9 * vk - 1e-4 * g
        assert np.allclose(p.detach().numpy(), wk, rtol=0, atol=1e-15)


def test_training_reduces_loss_on_a_small_buffer():
    from chinesecheckersagent_b200.train import TrainableResidualCNN, train
    torch.manual_seed(0)
    model = TrainableResidualCNN()
    gold = np.load(os.path.join(GOLDEN, "net_golden.npz"))
    planes = torch.from_numpy(gold["planes"][:160])
    pi = torch.zeros(160, 294); pi[torch.arange(160), torch.arange(160) % 294] = 1.0
    v = torch.sign(torch.randn(160))
    hist = train(model, planes, pi, v, epochs=3, seed=1)
    assert hist[-1]["policy"] < hist[0]["policy"] and all("val_loss" in h for h in hist)
