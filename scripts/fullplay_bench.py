"""cfg 5 played to completion (train.py:58-64 semantics): games/s with and without compaction of the draining batch."""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from chinesecheckersagent_b200.engine import Engine
from chinesecheckersagent_b200.model import ResidualCNN
from chinesecheckersagent_b200.selfplay import BatchedSelfPlay
eng = Engine(0)
m = ResidualCNN(engine=eng).load_weights(os.path.join(ROOT, 'tests', 'golden', 'good_model_weights.npz'))
games = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
for compact, ratio, poll in ((False, 0.5, 8), (True, 0.5, 8), (True, 0.75, 8), (True, 0.75, 4), (True, 0.85, 4)):
    sp = BatchedSelfPlay(eng, m.evaluate_states, n_slots=4096, seed=1, max_iters=512, ring=True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    st = sp.play_games(games, compact=compact, compact_ratio=ratio, poll_every=poll)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("compact=%s ratio %.2f poll %d: %d games in %.2f s = %.0f games/s, %d records, %d iterations, %d compactions, final batch %d" %
          (compact, ratio, poll, games, dt, games / dt, st["records"], st["iterations"], st["compactions"], sp.n), flush=True)
