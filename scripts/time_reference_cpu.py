"""Times the UNMODIFIED reference (/root/reference, through oracle/refshim.py) on this machine's CPU: BASELINE.md §3 items
1-3 (greedy-vs-greedy games, random-legal stepping, stub MCTS).  Build-container only (the GPU box has no reference
checkout); writes profiles/reference_cpu_container.json.  One core, like the reference's single-process scripts."""
import contextlib, io, json, os, platform, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import refshim
R = refshim.load()
import numpy as np
Board = R.board.Board
Game = __import__("game").Game
MCTSmod = __import__("MCTS")

out = {"cpu": platform.processor() or open("/proc/cpuinfo").read().split("model name")[1].split("\n")[0].strip(": \t"),
       "cores_used": 1, "os_cpu_count": os.cpu_count(), "python": sys.version.split()[0], "numpy": np.__version__}

# 1. Game('greedy','greedy').start()  (greedy_vs_greedy.py workload)
places = [0]
orig_place = Board.place
def counting_place(self, *a, **k):
    places[0] += 1
    return orig_place(self, *a, **k)
Board.place = counting_place
n_games, wins = 200, {1: 0, 2: 0, None: 0}
t = time.perf_counter()
with contextlib.redirect_stdout(io.StringIO()):
    for _ in range(n_games):
        wins[Game(p1_type='greedy', p2_type='greedy', verbose=False).start()] += 1
dt = time.perf_counter() - t
out["greedy_vs_greedy"] = {"games": n_games, "games_per_sec": n_games / dt, "plies_per_sec": places[0] / dt,
                           "plies_per_game": places[0] / n_games, "p1_wins": wins[1], "p2_wins": wins[2], "no_result": wins[None]}
Board.place = orig_place

# 2. random-legal stepping (selfplay.make_random_move's two-stage choice), ~5 s
steps, t = 0, time.perf_counter()
while time.perf_counter() - t < 5.0:
    b, player = Board(), 1
    for _ in range(200):
        vm = b.get_valid_moves(player)
        starts = [k for k in vm if vm[k]]
        s = random.choice(starts)
        if b.place(player, s, random.choice(vm[s])):
            break
        player = 3 - player
        steps += 1
out["random_env_steps_per_sec"] = steps / (time.perf_counter() - t)

# 3. MCTS with the uniform-prior stub, 175 simulations from the start position
class Stub:
    version = 0
    def predict(self, x):
        return np.full(294, 1 / 294.), 0.0
n_search, t = 3, time.perf_counter()
for _ in range(n_search):
    MCTSmod.MCTS(MCTSmod.Node(Board(), 1), Stub(), num_itr=175).search()
dt = time.perf_counter() - t
out["mcts_stub_sims_per_sec"] = n_search * 175 / dt
json.dump(out, open(os.path.join(ROOT, "profiles", "reference_cpu_container.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
