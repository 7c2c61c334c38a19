"""Second half of __graft_entry__.smoke(): one small pass through the tensor-core net, the fused net-MCTS rounds and the
greedy data generator, each checked against the CPU oracle (test infrastructure; imported here as the checker only)."""
import os

import numpy as np
import torch


def run(eng):
    import net_ref
    import oracle as orc
    from .data_generators import BatchedGreedyGenerator
    from .engine import BatchedMCTS
    from .model import ResidualCNN
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    wpath = os.path.join(root, "tests", "golden", "good_model_weights.npz")
    st, _, _ = orc.step_random(orc.start_states(96), 7, 0, 9)
    planes = orc.encode(st)
    model = ResidualCNN(engine=eng).load_weights(wpath)
    w = dict(np.load(wpath))
    p_ref, v_ref = net_ref.predict(w, planes, np.float64)
    x = torch.from_numpy(planes).to(eng.device)
    for kernel, bar in (("simt", 1e-4), ("tc", 5e-3)):
        model.set_kernel(kernel)
        p, v = model.predict_batch(x)
        dp = float(np.abs(p.cpu().numpy() - p_ref).max())
        dv = float(np.abs(v.cpu().numpy() - v_ref).max())
        assert dp < bar and dv < bar, "net (%s) differs from the restated graph: %g %g" % (kernel, dp, dv)
    model.set_kernel("tc")
    roots = torch.from_numpy(np.ascontiguousarray(st).view(np.int64)).to(eng.device)
    m = BatchedMCTS(eng, num_itr=24)
    a = m.search_net(roots)
    b = m.search_with(roots, model.evaluate_states)
    assert torch.equal(a["visits"], b["visits"]), "fused net-MCTS rounds differ from the round trips"
    assert int(a["visits"].sum(1).min()) == 23
    out = BatchedGreedyGenerator(eng, seed=5).generate(64)
    cand = orc.greedy_candidates(np.vstack([out["state"].cpu().numpy().view(np.uint64), np.zeros((3, out["state"].shape[1]), np.uint64)]))
    assert np.array_equal(cand, out["cand"].cpu().numpy().view(np.uint64)), "generator candidates differ from the oracle"
    return "+ net (simt, tcgen05) vs restated graph, fused net-MCTS rounds, greedy generator vs oracle"
