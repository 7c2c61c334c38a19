"""A/B of the greedy self-play kernel variants (CCX_GREEDY_VARIANT): bit-identity against variant 0 and timing (cfg 3)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from chinesecheckersagent_b200.engine import BatchedEnv, Engine
eng = Engine(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
ref = None
for variant in (0, 1, 2, 3):
    os.environ["CCX_GREEDY_VARIANT"] = str(variant)
    env = BatchedEnv(n, engine=eng, seed=0x5EED2026)
    c = env.play_greedy()
    st = env.state.clone(); cc = c.clone()
    ts = []
    for _ in range(5):
        env.reset()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); c2 = env.play_greedy(); b.record(); ts.append((a, b))
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in ts)[2]
    plies = int(cc[0].item())
    if ref is None:
        ref = (st, cc)
    same = bool(torch.equal(st, ref[0]) and torch.equal(cc, ref[1]))
    print("variant %d: %.3f ms  %.3e plies/s  %.3e games/s  identical=%s" % (variant, ms, plies / ms * 1e3, n / ms * 1e3, same), flush=True)
