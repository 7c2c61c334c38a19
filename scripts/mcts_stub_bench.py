"""cfg 4 stub search timing, repeated, to expose run-to-run variance."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from chinesecheckersagent_b200.engine import Engine, BatchedEnv, BatchedMCTS
eng = Engine(0)
env = BatchedEnv(4096, engine=eng); env.step_random(6)
m = BatchedMCTS(eng, num_itr=175)
m.search(env.state); torch.cuda.synchronize()
ts = []
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 20):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); m.search(env.state); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
print('search ms:', ' '.join('%.3f' % t for t in ts))
print('best %.3f median %.3f -> %.4g sims/s (median)' % (min(ts), sorted(ts)[len(ts) // 2], 4096 * 175 / sorted(ts)[len(ts) // 2] * 1e3))
