"""Phase timestamps of the trunk kernel's first tile (needs a build with CCX_NVCC_EXTRA=-DCCX_TRUNK_TIMING)."""
import ctypes, os, sys, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from chinesecheckersagent_b200.engine import Engine
from chinesecheckersagent_b200.model import ResidualCNN
eng = Engine(0)
m = ResidualCNN(engine=eng).load_weights(os.path.join(ROOT, 'tests', 'golden', 'good_model_weights.npz'))
x = torch.randint(0, 7, (4096, 7, 7, 7), dtype=torch.uint8, device='cuda')
for _ in range(3): m.forward(x)
torch.cuda.synchronize()
out = np.zeros(2048, dtype=np.int64)
lib = ctypes.CDLL(os.path.join(ROOT, 'chinesecheckersagent_b200', 'libccx.so'))
rc = lib.ccx_debug_trunk_timing(ctypes.c_void_p(out.ctypes.data))
ts = out.reshape(2, 1024)
for who in (0, 1):
    a = ts[who]; a = a[a > 0]
    d = np.diff(a)
    print('thread', 0 if who == 0 else 255, 'stamps', len(a), 'total cycles', a[-1] - a[0])
    # stamps cycle: layer_sync gives 4 stamps (0..3), wait_mma 2 (4,5): per phase 6 stamps
    k = (len(a) // 6) * 6
    b = a[:k].reshape(-1, 6)
    d6 = np.concatenate([np.diff(b, axis=1), np.append(b[1:, 0] - b[:-1, 5], 0)[:, None]], axis=1)
    names = ['cpwait', 'fence_async', 'sync', 'issue->waitstart', 'mma_wait', 'epilogue(to next layer_sync)']
    print('phases', d6.shape[0])
    for i, nme in enumerate(names):
        print('  %-30s mean %7.1f  median %7.1f  max %7d  sum %8d' % (nme, d6[:, i].mean(), np.median(d6[:, i]), d6[:, i].max(), d6[:, i].sum()))
    np.set_printoptions(linewidth=200)
    print(d6)
