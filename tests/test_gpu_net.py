"""GPU tests of the policy/value net kernels through the C-ABI against the float64 NumPy restatement.
Tolerance (BASELINE.json north_star): policy p and value v within 1e-3 abs; argmax agreement reported."""
import os

import numpy as np
import pytest
import torch

import net_ref
import oracle as orc
from conftest import GOLDEN

pytestmark = pytest.mark.gpu
WEIGHTS = os.path.join(GOLDEN, "good_model_weights.npz")


@pytest.fixture(scope="module")
def model():
    from chinesecheckersagent_b200.engine import Engine
    from chinesecheckersagent_b200.model import ResidualCNN
    m = ResidualCNN(engine=Engine(0)).load_weights(WEIGHTS)
    yield m
    m.eng.close()


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN, "net_golden.npz"))


@pytest.mark.parametrize("dtype", [torch.uint8, torch.float32, torch.bfloat16])
def test_forward_matches_restatement(model, gold, dtype):
    planes = torch.from_numpy(gold["planes"]).cuda().to(dtype)
    logits, value = model.forward(planes)
    logits, value = logits.cpu().numpy().astype(np.float64), value.cpu().numpy().astype(np.float64)
    assert np.abs(logits - gold["logits"]).max() < 1e-3          # fp32 SIMT kernel: ~1e-5 in practice
    assert np.abs(value - gold["v"]).max() < 1e-3
    p_ref = net_ref.softmax64(gold["logits"])
    p, v = model.predict_batch(planes)
    assert np.abs(p.cpu().numpy() - p_ref).max() < 1e-3
    assert np.abs(v.cpu().numpy() - gold["v"]).max() < 1e-3
    assert (p.cpu().numpy().argmax(1) == p_ref.argmax(1)).mean() >= 0.99     # move-agreement rate
    assert np.allclose(p.cpu().numpy().sum(1), 1.0, atol=1e-12)


def test_predict_single_position_like_reference_api(model):
    x = orc.encode(orc.start_states(1))[0].astype(np.float64)                 # utils.to_model_input output
    p, v = model.predict(x)
    assert p.shape == (294,) and p.dtype == np.float64
    assert abs(float(v) - (-0.043674)) < 1e-4                                  # SURVEY.md §8c known answer
    assert list(np.argsort(-p)[:5]) == [76, 117, 101, 60, 143]


@pytest.mark.parametrize("n", [1, 7, 8, 9, 1000])
def test_ragged_batches(model, gold, n):
    idx = np.arange(n) % len(gold["planes"])
    planes = torch.from_numpy(gold["planes"][idx]).cuda()
    logits, value = model.forward(planes)
    assert np.abs(logits.cpu().numpy() - gold["logits"][idx]).max() < 1e-3
    assert np.abs(value.cpu().numpy() - gold["v"][idx]).max() < 1e-3


def test_evaluate_states_fuses_encode_and_predict(model):
    st, _, _ = orc.step_random(orc.start_states(300), 5, 0, 11)
    dev = torch.from_numpy(np.ascontiguousarray(st[:5]).view(np.int64)).cuda()
    p, v = model.evaluate_states(dev)
    w = dict(np.load(WEIGHTS))
    p_ref, v_ref = net_ref.predict(w, orc.encode(st), np.float64)
    assert np.abs(p.cpu().numpy() - p_ref).max() < 1e-3 and np.abs(v.cpu().numpy() - v_ref).max() < 1e-3


def test_mcts_with_net_runs_and_is_deterministic(model):
    from chinesecheckersagent_b200.engine import BatchedMCTS
    st, _, _ = orc.step_random(orc.start_states(64), 5, 0, 6)
    roots = torch.from_numpy(np.ascontiguousarray(st).view(np.int64)).cuda()
    m = BatchedMCTS(model.eng, num_itr=40)
    a = m.search_with(roots, model.evaluate_states)
    b = m.search_with(roots, model.evaluate_states)
    assert torch.equal(a["visits"], b["visits"])
    assert np.all(a["visits"].cpu().numpy().sum(1) == 39)
