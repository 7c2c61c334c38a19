"""Phase timestamps of the accurate trunk (build with -DCCX_ACC_TIMING into chinesecheckersagent_b200/libccx_timing.so and run with
CCX_LIB_PATH pointing at it): prints clock64 deltas of block 0 / thread 0 between the stamps of one tile (see ACC_TS in ccx_net_tc.cu)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from chinesecheckersagent_b200.engine import Engine
from chinesecheckersagent_b200.model import ResidualCNN
m = ResidualCNN(engine=Engine(0)).load_weights(os.path.join(ROOT, 'tests', 'golden', 'good_model_weights.npz'))
m.set_kernel('tc_acc')
ctx, n = int(sys.argv[1]), int(sys.argv[2])
m.eng.call('ccx_net_set_acc_contexts', ctx)
x = torch.randint(0, 7, (n, 7, 7, 7), dtype=torch.uint8, device='cuda')
for _ in range(2):
    m.forward(x)
    torch.cuda.synchronize()
    print('---', flush=True)
