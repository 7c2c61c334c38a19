"""A/B of the env step kernel variants (CCX_STEP_VARIANT, CCX_STEP_THRESH): bit-identity against variant 0 and timing.
usage: python scripts/env_variants.py [games ...]"""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from chinesecheckersagent_b200.engine import BatchedEnv, Engine

eng = Engine(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
sizes = [int(x) for x in sys.argv[1:]] or [65536, 1048576]


def run(n, variant, thresh=None, reps=5):
    os.environ["CCX_STEP_VARIANT"] = str(variant)
    if thresh is None:
        os.environ.pop("CCX_STEP_THRESH", None)
    else:
        os.environ["CCX_STEP_THRESH"] = str(thresh)
    env = BatchedEnv(n, engine=eng, seed=0x5EED2026)
    env.step_random(256)
    final = env.state.clone(); wins = env.wins.clone()
    for _ in range(2):
        env.step_random(256)
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        flush.fill_(1); a.record(); env.step_random(256); b.record(); ts.append((a, b))
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in ts)[len(ts) // 2]
    return final, wins, ms


for n in sizes:
    ref, refw, ms0 = run(n, 0)
    print("n=%d variant 0 (flat, 1 game/lane): %.3f ms  %.3e steps/s" % (n, ms0, n * 256 / ms0 * 1e3), flush=True)
    for variant, threshes in ((9, [None]),):
        for th in threshes:
            st, w, ms = run(n, variant, th)
            same = bool(torch.equal(st[:5], ref[:5]) and torch.equal(w, refw))
            print("n=%d variant %d thresh %s: %.3f ms  %.3e steps/s  (%.2fx)  identical=%s" % (n, variant, th, ms, n * 256 / ms * 1e3, ms0 / ms, same), flush=True)
