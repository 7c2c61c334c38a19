"""Latency of the drop-in one-tree search, MCTS(Node(Board()), model).search() (player.py:157-158), per net mode."""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from chinesecheckersagent_b200.engine import Engine
from chinesecheckersagent_b200.model import ResidualCNN
from chinesecheckersagent_b200.board import Board
from chinesecheckersagent_b200.MCTS import MCTS, Node
eng = Engine(0)
m = ResidualCNN(engine=eng).load_weights(os.path.join(ROOT, 'tests', 'golden', 'good_model_weights.npz'))
for kern in ("tc_acc", "tc"):
    m.set_kernel(kern)
    f = lambda: MCTS(Node(Board(engine=eng), 1), m, num_itr=175).search()
    for _ in range(3): f()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): f()
    torch.cuda.synchronize()
    print("%s: %.3f ms per 175-simulation decision" % (kern, (time.perf_counter() - t0) / 10 * 1e3))
