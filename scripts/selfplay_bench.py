"""cfg 5 self-play step timing (4,096 slots, 175 sims + root expansion per ply); used under ncu for launch lists."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from chinesecheckersagent_b200.engine import Engine
from chinesecheckersagent_b200.model import ResidualCNN
from chinesecheckersagent_b200.selfplay import BatchedSelfPlay
eng = Engine(0)
m = ResidualCNN(engine=eng).load_weights(os.path.join(ROOT, 'tests', 'golden', 'good_model_weights.npz'))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
kernel = sys.argv[3] if len(sys.argv) > 3 else 'tc'
m.set_kernel(kernel)
sp = BatchedSelfPlay(eng, m.evaluate_states, n_slots=n, max_iters=32)
for _ in range(7): sp.step()
torch.cuda.synchronize()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(reps): sp.step()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / reps
print('selfplay ply (%s): %.3f ms -> %.4g sims/s (%d slots)' % (kernel, ms, n * 175 / ms * 1e3, n))
