"""Small end-to-end pass over every kernel family for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from chinesecheckersagent_b200.engine import Engine, BatchedEnv, BatchedMCTS
from chinesecheckersagent_b200.model import ResidualCNN
from chinesecheckersagent_b200.selfplay import BatchedSelfPlay
from chinesecheckersagent_b200.data_generators import BatchedGreedyGenerator
from chinesecheckersagent_b200.arena import BatchedArena, GREEDY
eng = Engine(0)
env = BatchedEnv(300, engine=eng)
env.step_random(24); env.movegen(); env.encode(); env.greedy_candidates()
g = BatchedEnv(200, engine=eng); g.play_greedy()
BatchedMCTS(eng, num_itr=12).search(env.state)
m = ResidualCNN(engine=eng).load_weights(os.path.join(ROOT, 'tests', 'golden', 'good_model_weights.npz'))
for n in (3, 130):
    x = torch.randint(0, 7, (n, 7, 7, 7), dtype=torch.uint8, device='cuda')
    m.set_kernel('tc'); m.forward(x); m.set_kernel('tc_acc'); m.forward(x); m.set_kernel('simt'); m.forward(x)
m.set_kernel('tc')
BatchedMCTS(eng, num_itr=6).search_net(env.state[:, :70].contiguous())
g9 = BatchedMCTS(eng, num_itr=9); r9 = env.state[:, :70].contiguous()
for _ in range(3): g9.search_net(r9)          # direct, captured, replayed (CUDA graph of the round loop)
m.set_kernel('tc_acc')
for _ in range(3): g9.search_net(r9)
m.set_kernel('tc')
sp = BatchedSelfPlay(eng, m.evaluate_states, n_slots=33, num_itr=5, max_iters=10)
for _ in range(8): sp.step()
sp.collect()
BatchedGreedyGenerator(eng).generate(50, random_start=True)
a = BatchedArena(m, GREEDY, 9, num_itr=4)
for _ in range(4): a.step()
torch.cuda.synchronize()
print('sanitize pass done')
