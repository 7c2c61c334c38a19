"""Runs the UNMODIFIED Python reference (kenziyuliu/ChineseCheckersAgent) on the host cores — TEST / BASELINE
INFRASTRUCTURE, NOT PRODUCT.

Importers: bench.py's `cpu_baseline` / `--impl reference` legs and tests/.  The product package never imports this.

The reference is a flat collection of .py files with no native code, so there is nothing to compile into
`oracle/_ref/`: the "recipe" is `stage()` below, which packs the .py files of the read-only checkout
(`/root/reference`, build container only) into ONE archive, `oracle/_ref/reference.zip`, which Python imports from
directly (zipimport).  `oracle/_ref/` is git-ignored (the sources never enter this repository, neither its history
nor its tree of source files) but travels to the GPU box with the gpurun snapshot like a built `.so`, so the
reference's own `Board` / `Game` / `MCTS` can be timed on the box's host cores in the same run as the GPU numbers
(BASELINE.md §3).  Every function here resolves the reference through `refshim` (stubs for the missing
h5py / keras / tensorflow imports; the Keras net is never instantiated).

Workloads:
  random_steps   cfg 2: per game `Board()`, then per ply `get_valid_moves` + selfplay.make_random_move's two-stage
                 choice (selfplay.py:93-98) + `Board.place` (board.py:226-250, returns check_win); a won game restarts
  greedy_games   cfg 1: `Game('greedy','greedy',verbose=False).start()` (greedy_vs_greedy.py:17-19, game.py:58-100)
  mcts_stub      cfg 4: `MCTS(Node(Board(),1), StubModel(), num_itr=175).search()` with the uniform-prior stub
"""
import contextlib
import glob
import io
import os
import random
import sys
import time
import zipfile

_HERE = os.path.dirname(os.path.abspath(__file__))
STAGED = os.path.join(_HERE, "_ref", "reference.zip")
SOURCE = "/root/reference"


def stage(force=False):
    """Pack the reference's .py files into oracle/_ref/reference.zip (only where the checkout exists).  Returns the archive
    path, or None when neither the checkout nor an earlier staging is available."""
    if os.path.isdir(SOURCE):
        srcs = sorted(glob.glob(os.path.join(SOURCE, "*.py")))
        newest = max(os.path.getmtime(s) for s in srcs)
        if force or not os.path.exists(STAGED) or os.path.getmtime(STAGED) < newest:
            os.makedirs(os.path.dirname(STAGED), exist_ok=True)
            tmp = STAGED + ".tmp"
            with zipfile.ZipFile(tmp, "w", zipfile.ZIP_DEFLATED) as z:
                for s in srcs:
                    z.write(s, os.path.basename(s))
                z.writestr("STAGED_FROM", "%s (%d .py files, packed by oracle/refrun.py:stage; git-ignored, travels with gpurun)\n"
                           % (SOURCE, len(srcs)))
            os.replace(tmp, STAGED)
    return STAGED if os.path.exists(STAGED) else None


def reference_dir():
    """The staged archive when present (it is what travels), else the read-only checkout, else None.  Either one works as a
    sys.path entry."""
    if os.path.exists(STAGED):
        return STAGED
    if os.path.exists(os.path.join(SOURCE, "board.py")):
        return SOURCE
    return None


def reference_source(name):
    """text of one reference file (from the archive or the checkout)"""
    d = reference_dir()
    if d is None:
        raise RuntimeError("no reference available")
    if d.endswith(".zip"):
        with zipfile.ZipFile(d) as z:
            return z.read(name).decode()
    with open(os.path.join(d, name)) as f:
        return f.read()


def available():
    return reference_dir() is not None


def _load():
    d = reference_dir()
    if d is None:
        raise RuntimeError("the reference is neither staged as oracle/_ref/reference.zip nor present at /root/reference")
    os.environ["CCX_REFERENCE_DIR"] = d
    if _HERE not in sys.path:
        sys.path.insert(0, _HERE)
    import refshim
    refshim.REFERENCE_DIR = d
    return refshim.load()


# ---- workers (module-level so that multiprocessing can pickle them) ---------------------------------------------

def _random_steps_worker(args):
    games, plies, seed = args
    R = _load()
    Board = R.board.Board
    rng = random.Random(seed)
    steps = wins = 0
    t0 = time.perf_counter()
    for _ in range(games):
        b, player = Board(), 1
        for _ in range(plies):
            vm = b.get_valid_moves(player)                        # board.py:215-222
            keys = list(vm.keys())
            s = rng.choice(keys)                                  # selfplay.py:95-97
            while len(vm[s]) == 0:
                s = rng.choice(keys)
            won = b.place(player, s, rng.choice(vm[s]))           # selfplay.py:98,101; board.py:226-250
            steps += 1
            if won:
                wins += 1
                b, player = Board(), 1
            else:
                player = 3 - player
    return steps, wins, time.perf_counter() - t0


def _greedy_games_worker(args):
    games, seed = args
    R = _load()
    random.seed(seed)
    Board, Game = R.board.Board, R.game.Game
    places = [0]
    orig = Board.place

    def counting_place(self, *a, **k):
        places[0] += 1
        return orig(self, *a, **k)
    Board.place = counting_place
    wins = {1: 0, 2: 0, None: 0}
    t0 = time.perf_counter()
    try:
        with contextlib.redirect_stdout(io.StringIO()):           # Game.start prints unconditionally (game.py:80,98)
            for _ in range(games):
                wins[Game(p1_type='greedy', p2_type='greedy', verbose=False).start()] += 1
    finally:
        Board.place = orig
    return games, places[0], wins[1], wins[2], wins[None], time.perf_counter() - t0


class _Stub:
    version = 0

    def predict(self, x):
        import numpy as np
        return np.full(294, 1 / 294.), 0.0


def _mcts_stub_worker(args):
    searches, sims = args
    R = _load()
    t0 = time.perf_counter()
    for _ in range(searches):
        R.MCTS.MCTS(R.MCTS.Node(R.board.Board(), 1), _Stub(), num_itr=sims).search()
    return searches * sims, time.perf_counter() - t0


# ---- drivers ----------------------------------------------------------------------------------------------------

class Pool:
    """multiprocessing.Pool(os.cpu_count()) of reference workers, kept alive across bench steps (the reference's own
    parallelism is a process pool too, train.py:71-86)."""

    def __init__(self, procs=None, spawn=False):
        import multiprocessing as mp
        self.procs = int(procs or os.cpu_count() or 1)
        # spawn = True when the parent has a CUDA context (bench.py's GPU arm): forked children of such a process are unsafe
        self.pool = mp.get_context("spawn" if spawn else "fork").Pool(self.procs) if self.procs > 1 else None

    def map(self, fn, jobs):
        return self.pool.map(fn, jobs) if self.pool else [fn(j) for j in jobs]

    def close(self):
        if self.pool:
            self.pool.close()
            self.pool.join()
            self.pool = None

    def random_steps(self, games, plies, seed=0):
        """`games` games of `plies` random-legal plies spread over the pool -> (env steps, wins, wall seconds)"""
        per = [games // self.procs + (1 if i < games % self.procs else 0) for i in range(self.procs)]
        t0 = time.perf_counter()
        res = self.map(_random_steps_worker, [(g, plies, seed * 1000003 + i) for i, g in enumerate(per) if g])
        return sum(r[0] for r in res), sum(r[1] for r in res), time.perf_counter() - t0

    def greedy_games(self, games, seed=0):
        per = [games // self.procs + (1 if i < games % self.procs else 0) for i in range(self.procs)]
        t0 = time.perf_counter()
        res = self.map(_greedy_games_worker, [(g, seed * 1000003 + i) for i, g in enumerate(per) if g])
        dt = time.perf_counter() - t0
        return dict(games=sum(r[0] for r in res), plies=sum(r[1] for r in res), p1_wins=sum(r[2] for r in res),
                    p2_wins=sum(r[3] for r in res), no_result=sum(r[4] for r in res), seconds=dt)

    def mcts_stub(self, searches, sims=175):
        per = [searches // self.procs + (1 if i < searches % self.procs else 0) for i in range(self.procs)]
        t0 = time.perf_counter()
        res = self.map(_mcts_stub_worker, [(s, sims) for s in per if s])
        return sum(r[0] for r in res), time.perf_counter() - t0


def greedy_vs_greedy_script():
    """greedy_vs_greedy.py as shipped (50 games, `python greedy_vs_greedy.py`), run through the shim in this process.
    Returns (seconds, captured tail of its output)."""
    _load()
    code = compile(reference_source("greedy_vs_greedy.py"), "greedy_vs_greedy.py", "exec")
    buf = io.StringIO()
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(buf):
        exec(code, {"__name__": "__main__"})
    dt = time.perf_counter() - t0
    lines = [l for l in buf.getvalue().splitlines() if l.strip()]
    return dt, lines[-3:]


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


if __name__ == "__main__":
    print(stage(force="--force" in sys.argv))
