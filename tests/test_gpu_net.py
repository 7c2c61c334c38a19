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
    model.set_kernel("simt")
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


@pytest.mark.parametrize("tc_dtype,bar_p,bar_v", [("fp16", 5e-3, 4e-3), ("bf16", 5e-2, 2e-2)])
def test_16bit_throughput_mode_regression_bounds(model, gold, tc_dtype, bar_p, bar_v):
    """NOT a parity test: the 16-bit tcgen05 mode (`set_kernel('tc')`) is OUT of the north_star's 1e-3 tolerance and is never the
    default; parity of the net is asserted on the accurate mode and the fp32 kernel
    (test_accurate_tensor_core_mode_meets_the_1e3_bar, test_forward_matches_restatement, test_net_bar_on_10k_positions_...).
    This test only keeps the throughput mode from regressing: its error stays at the level measured on B200 (256 fixture
    positions: fp16 max|dp| 3.2e-3, mean 5.5e-6, max|dv| 2.0e-3; bf16 max|dp| 2.8e-2, max|dv| 9.6e-3; ~1.6x margin) and its
    arg-max move agrees with the reference's."""
    planes = torch.from_numpy(gold["planes"]).cuda()
    model.set_kernel("tc", tc_dtype=tc_dtype)
    ltc, vtc = model.forward(planes)
    p_ref = net_ref.softmax64(gold["logits"])
    p_tc = net_ref.softmax64(ltc.cpu().numpy())
    dp = np.abs(p_tc - p_ref).max()
    dv = np.abs(vtc.cpu().numpy() - gold["v"]).max()
    agree = (p_tc.argmax(1) == p_ref.argmax(1)).mean()
    print("tc %s: max|dp| %.4g max|dv| %.4g argmax agreement %.4f" % (tc_dtype, dp, dv, agree))
    assert dp < bar_p and dv < bar_v and agree >= 0.99
    model.set_kernel("tc", tc_dtype="fp16")


@pytest.mark.parametrize("n", [1, 4, 5, 6, 127, 128, 129, 1000])
def test_tensor_core_kernel_ragged_batches(model, gold, n):
    model.set_kernel("tc")
    idx = np.arange(n) % len(gold["planes"])
    planes = torch.from_numpy(gold["planes"][idx]).cuda()
    logits, value = model.forward(planes)
    assert np.abs(logits.cpu().numpy() - gold["logits"][idx]).max() < 0.25
    assert np.abs(value.cpu().numpy() - gold["v"][idx]).max() < 2e-2


def test_predict_single_position_like_reference_api(model):
    model.set_kernel("simt")
    x = orc.encode(orc.start_states(1))[0].astype(np.float64)                 # utils.to_model_input output
    p, v = model.predict(x)
    assert p.shape == (294,) and p.dtype == np.float64
    assert abs(float(v) - (-0.043674)) < 1e-4                                  # SURVEY.md §8c known answer
    assert list(np.argsort(-p)[:5]) == [76, 117, 101, 60, 143]


@pytest.mark.parametrize("n", [1, 7, 8, 9, 1000])
def test_ragged_batches(model, gold, n):
    model.set_kernel("simt")
    idx = np.arange(n) % len(gold["planes"])
    planes = torch.from_numpy(gold["planes"][idx]).cuda()
    logits, value = model.forward(planes)
    assert np.abs(logits.cpu().numpy() - gold["logits"][idx]).max() < 1e-3
    assert np.abs(value.cpu().numpy() - gold["v"][idx]).max() < 1e-3


def test_evaluate_states_fuses_encode_and_predict(model):
    model.set_kernel("simt")
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


@pytest.mark.parametrize("kernel", ["simt", "tc"])
@pytest.mark.parametrize("pre_expand", [False, True])
def test_fused_round_loop_equals_round_trips(model, kernel, pre_expand):
    """ccx_mcts_run_net (select+encode, net, softmax+expand+backup fused, all rounds in one C call) must build
    exactly the trees of the select / ccx_net_eval / expand_backup round trips: same visit counts, same root Q."""
    from chinesecheckersagent_b200.engine import BatchedMCTS
    model.set_kernel(kernel)
    st, _, _ = orc.step_random(orc.start_states(257), 11, 0, 9)
    roots = torch.from_numpy(np.ascontiguousarray(st).view(np.int64)).cuda()
    noise = torch.rand((257, 128), dtype=torch.float64, device="cuda") if pre_expand else None
    m = BatchedMCTS(model.eng, num_itr=48)
    a = m.search_with(roots, model.evaluate_states, pre_expand=pre_expand, root_noise=noise)
    b = m.search_net(roots, pre_expand=pre_expand, root_noise=noise)
    assert torch.equal(a["visits"], b["visits"])
    assert torch.equal(a["q"].view(torch.int64), b["q"].view(torch.int64))
    assert torch.equal(a["n_nodes"], b["n_nodes"])
    model.set_kernel("tc")


@pytest.mark.parametrize("kernel", ["tc", "tc_acc"])
def test_fused_round_loop_two_stream_split_is_bit_identical(model, kernel):
    """>= 8,192 trees (16,384 in the accurate mode): ccx_mcts_run_net runs the two halves of the batch as two pipelines on two streams (ragged halves here);
    the trees must not depend on that (16-bit and accurate tensor-core modes; rounds 10 >= 8, so the third call also
    replays the two-branch CUDA graph)"""
    from chinesecheckersagent_b200.engine import BatchedMCTS
    model.set_kernel(kernel)
    n = (16400 if kernel == "tc_acc" else 8200) + 3
    st, _, _ = orc.step_random(orc.start_states(n), 21, 0, 7, nthreads=8)
    roots = torch.from_numpy(np.ascontiguousarray(st).view(np.int64)).cuda()
    noise = torch.rand((n, 128), dtype=torch.float64, device="cuda")
    m = BatchedMCTS(model.eng, num_itr=9)
    a = m.search_with(roots, model.evaluate_states, pre_expand=True, root_noise=noise)
    b = m.search_net(roots, pre_expand=True, root_noise=noise)
    assert torch.equal(a["visits"], b["visits"]) and torch.equal(a["q"].view(torch.int64), b["q"].view(torch.int64))
    assert torch.equal(a["n_nodes"], b["n_nodes"])
    for _ in range(2):                                                   # captured, then replayed
        c = m.search_net(roots, pre_expand=True, root_noise=noise)
        assert torch.equal(a["visits"], c["visits"]) and torch.equal(a["n_nodes"], c["n_nodes"])
    model.set_kernel("tc")


def test_accurate_tensor_core_mode_meets_the_1e3_bar(model, gold):
    """split-precision tcgen05 path (hi + lo halves of activations and weights, fp32 policy dense): the north_star's
    |dp|, |dv| <= 1e-3 against the float64 restatement, with margin; argmax agreement 100 %"""
    planes = torch.from_numpy(gold["planes"]).cuda()
    model.set_kernel("tc_acc")
    l, v = model.forward(planes)
    p_ref = net_ref.softmax64(gold["logits"])
    p = net_ref.softmax64(l.cpu().numpy())
    dp = np.abs(p - p_ref).max()
    dv = np.abs(v.cpu().numpy() - gold["v"]).max()
    dl = np.abs(l.cpu().numpy() - gold["logits"]).max()
    print("tc_acc: max|dlogit| %.4g max|dp| %.4g max|dv| %.4g" % (dl, dp, dv))
    assert dp < 1e-4 and dv < 1e-4 and (p.argmax(1) == p_ref.argmax(1)).all()
    for n in (1, 3, 4, 5, 130):                                      # ragged batches
        l2, v2 = model.forward(planes[:n])
        assert torch.equal(l2, l[:n]) and torch.equal(v2, v[:n])
    model.set_kernel("tc")


@pytest.mark.parametrize("n", [1, 4, 5, 130, 592, 1777, 4096, 7001])
def test_acc_multi_context_equals_single(model, n):
    """k_net_trunk_accm<1..3> (one CTA per SM, 1-3 tiles in flight sharing the weight slots) computes k_net_trunk_acc's bits:
    batches that leave contexts without a tile, fill exactly one round, and run several rounds with a partial last one"""
    g = torch.Generator().manual_seed(n)
    planes = torch.randint(0, 4, (n, 7, 7, 7), dtype=torch.uint8, generator=g).cuda()
    model.set_kernel("tc_acc")
    try:
        model.eng.call("ccx_net_set_acc_contexts", 0)
        l0, v0 = model.forward(planes)
        l0, v0 = l0.clone(), v0.clone()
        for ctx in (1, 2, 3):
            model.eng.call("ccx_net_set_acc_contexts", ctx)
            for _ in range(2):                                   # twice: the second launch finds the slots' barriers re-initialised
                l, v = model.forward(planes)
                assert torch.equal(l, l0) and torch.equal(v, v0), "contexts=%d differs" % ctx
    finally:
        model.eng.call("ccx_net_set_acc_contexts", -1)
        l, v = model.forward(planes)                             # automatic choice
        assert torch.equal(l, l0) and torch.equal(v, v0)
        model.set_kernel("tc")


def test_accurate_mode_in_the_fused_mcts_rounds(model):
    """ccx_mcts_run_net with the accurate net = the select / ccx_net_eval / expand_backup round trips in the same mode"""
    from chinesecheckersagent_b200.engine import BatchedMCTS
    model.set_kernel("tc_acc")
    st, _, _ = orc.step_random(orc.start_states(200), 31, 0, 8)
    roots = torch.from_numpy(np.ascontiguousarray(st).view(np.int64)).cuda()
    m = BatchedMCTS(model.eng, num_itr=30)
    a = m.search_with(roots, model.evaluate_states)
    b = m.search_net(roots)
    assert torch.equal(a["visits"], b["visits"]) and torch.equal(a["q"].view(torch.int64), b["q"].view(torch.int64))
    model.set_kernel("tc")


def test_net_bar_on_10k_positions_met_in_self_play(model):
    """SURVEY 8d cfg 5: net parity on >= 10,000 positions ENCOUNTERED in self-play (searched plies of 1,024 games driven by
    the net itself), against a float64 evaluation of the same Keras graph (train.TrainableResidualCNN in eval mode, which
    tests/test_train_cpu.py pins to the NumPy restatement at 1e-9).  Bars: |dp|, |dv| <= 1e-3 (north_star) for the accurate
    default mode and the fp32 kernel, argmax-move agreement reported and >= 99.9 %; the 16-bit throughput mode is reported."""
    from chinesecheckersagent_b200.model import read_weight_file
    from chinesecheckersagent_b200.selfplay import BatchedSelfPlay
    from chinesecheckersagent_b200.train import TrainableResidualCNN
    model.set_kernel("tc_acc")
    sp = BatchedSelfPlay(model.eng, model.evaluate_states, n_slots=1024, num_itr=12, max_iters=24, seed=77)
    for _ in range(6 + 12):
        sp.step()
    flags = sp.rec_flag & 0xF                                           # every searched ply, finished game or not
    rows = torch.nonzero(flags != 0).flatten()
    assert rows.numel() >= 10000
    state = sp.rec_state[rows].t().contiguous()
    planes = model.eng.empty((rows.numel(), 7, 7, 7), torch.uint8)
    from chinesecheckersagent_b200.config import DTYPE_U8
    from chinesecheckersagent_b200.engine import _p
    model.eng.call("ccx_encode", rows.numel(), _p(state), _p(planes), DTYPE_U8)
    assert len(torch.unique(state.t(), dim=0)) >= 9000                  # genuinely different positions
    ref = TrainableResidualCNN().load_keras_weights(read_weight_file(WEIGHTS)).double().cuda().eval()
    with torch.no_grad():
        logits, v_ref = ref(planes)
    p_ref = torch.softmax(logits, dim=1)
    for kernel, bar in (("tc_acc", 1e-3), ("simt", 1e-3), ("tc", None)):
        model.set_kernel(kernel)
        p, v = model.predict_batch(planes)
        dp, dv = (p - p_ref).abs().max().item(), (v - v_ref).abs().max().item()
        agree = (p.argmax(1) == p_ref.argmax(1)).double().mean().item()
        print("%s on %d self-play positions: max|dp| %.3g max|dv| %.3g argmax agreement %.5f" % (kernel, rows.numel(), dp, dv, agree))
        if bar is not None:
            assert dp <= bar and dv <= bar and agree >= 0.999, kernel
        else:
            assert dp < 2e-2 and dv < 2e-2 and agree >= 0.99
    model.set_kernel("tc")


def test_round_loop_graph_replay_is_identical(model):
    """ccx_mcts_run_net launches a cached CUDA graph of the round loop from the second call with the same arguments on; every
    call (direct, captured, replayed) must equal the unfused select / evaluate / expand_backup round trips, and a change of
    evaluator mode, batch size or weights must drop the cached graph."""
    from chinesecheckersagent_b200.engine import BatchedMCTS
    if os.environ.get("CCX_NO_GRAPH"):
        pytest.skip("graph replay switched off")
    L, h = model.eng.L, model.eng.h
    st, _, _ = orc.step_random(orc.start_states(200), 31, 0, 8)
    big = torch.from_numpy(np.ascontiguousarray(st).view(np.int64)).cuda()
    small = big[:, :70].contiguous()
    m = BatchedMCTS(model.eng, num_itr=20)
    replays0 = L.ccx_graph_replays(h)
    for kernel in ("tc", "tc_acc", "tc", "simt"):
        model.set_kernel(kernel)
        for roots in (big, small, big):
            want = m.search_with(roots, model.evaluate_states)
            before = L.ccx_graph_replays(h)
            for rep in range(4):
                got = m.search_net(roots)
                assert torch.equal(want["visits"], got["visits"]), (kernel, roots.shape, rep)
                assert torch.equal(want["q"].view(torch.int64), got["q"].view(torch.int64))
            assert L.ccx_graph_replays(h) - before == 3                 # call 1 ran directly, call 2 captured + launched, 3 and 4 replayed
    model.load_weights(WEIGHTS)                                         # reload: the graph of the old buffers is dropped
    model.set_kernel("tc")
    want = m.search_with(big, model.evaluate_states)
    before = L.ccx_graph_replays(h)
    for rep in range(3):
        assert torch.equal(want["visits"], m.search_net(big)["visits"])
    assert L.ccx_graph_replays(h) - before == 2
    assert L.ccx_graph_replays(h) > replays0
