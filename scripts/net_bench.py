"""Net forward timing at the self-play batch (4096) and a large batch; used under ncu for kernel captures."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from chinesecheckersagent_b200.engine import Engine
from chinesecheckersagent_b200.model import ResidualCNN
m = ResidualCNN(engine=Engine(0)).load_weights(os.path.join(ROOT, 'tests', 'golden', 'good_model_weights.npz'))
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
kernel = sys.argv[2] if len(sys.argv) > 2 else 'tc'
m.set_kernel(kernel)
if len(sys.argv) > 3: m.eng.call('ccx_net_set_acc_contexts', int(sys.argv[3]))      # accurate trunk: tiles in flight per CTA (0 = one tile per CTA)
for n in (4096, 65536):
    x = torch.randint(0, 7, (n, 7, 7, 7), dtype=torch.uint8, device='cuda')
    for _ in range(3): m.forward(x)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): m.forward(x)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    print(kernel + (' ctx' + sys.argv[3] if len(sys.argv) > 3 else '') + ' forward %d positions: %.4f ms -> %.4g positions/s, %.1f TFLOP/s' % (n, ms, n / ms * 1e3, 6.483264e6 * n / ms / 1e9))
