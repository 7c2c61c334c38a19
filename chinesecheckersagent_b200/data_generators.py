"""Mirror of the reference's data_generators.py: `GreedyDataGenerator(randomised, random_start).generate_play()`
-> (play_history [(Board, pi[294])], reward for player 1), and the batched generator that is the product:
`BatchedGreedyGenerator(engine).generate(n, ...)` -> board_x / pi_y / v_y in utils.convert_to_train_data's format
(utils.py:60-73), produced on the device at ~1e8 games/s (csrc/ccx_datagen.cu).

Deviation flagged (SURVEY §8a quirk iii): utils.convert_to_train_data assumes the first kept record is player 1's,
which is wrong for randomised games (three dropped plies); records here always carry the true side to move, so
board_x and v_y of randomised games are the correct ones, not the reference's mislabelled ones."""
import torch

from .config import AVERAGE_TOTAL_MOVE, BOARD_HIST_MOVES, DEFAULT_SEED, DTYPE_U8, INITIAL_RANDOM_MOVES, STUCK_PLY_LIMIT
from .engine import BatchedEnv, _p


class BatchedGreedyGenerator:
    def __init__(self, engine, seed=DEFAULT_SEED, rank=0, world=1):
        self.eng, self.seed, self.rank, self.world = engine, int(seed), int(rank), int(world)
        self.games_done = 0

    def generate(self, n, randomised=False, random_start=False, stuck_plies=STUCK_PLY_LIMIT, want_pi=True, want_planes=True):
        """n games of one kind.  Returns dict(board_x (M,7,7,7) uint8, pi_y (M,294) float32, v_y (M,) int8,
        state (5,M) int64, cand (6,M) int64, game (M,) int32, lengths (n,) int32, winners (n,) uint8)."""
        e = self.eng
        gid0 = self.games_done * self.world + self.rank * n          # disjoint global game ids across calls and ranks
        self.games_done += n
        env = BatchedEnv(n, engine=e, seed=self.seed, game_id0=gid0, randomised=randomised)
        lengths = e.empty((n,), torch.int32)
        winners = e.empty((n,), torch.uint8)
        args = (n, _p(env.state), gid0, self.seed, INITIAL_RANDOM_MOVES if random_start else 0,
                BOARD_HIST_MOVES if randomised else 0, int(stuck_plies), AVERAGE_TOTAL_MOVE)
        e.call("ccx_greedy_generate", *args, _p(lengths), _p(winners), None, 0, None, None, None, None)
        offsets = (torch.cumsum(lengths, 0, dtype=torch.int64) - lengths).contiguous()
        m = int(lengths.sum().item())
        state = e.empty((5, max(m, 1)), torch.int64)
        cand = e.empty((6, max(m, 1)), torch.int64)
        v_y = e.empty((max(m, 1),), torch.int8)
        game = e.empty((max(m, 1),), torch.int32)
        e.call("ccx_greedy_generate", *args, _p(lengths), _p(winners), _p(offsets), m, _p(state), _p(cand), _p(v_y), _p(game))
        out = dict(state=state[:, :m], cand=cand[:, :m], v_y=v_y[:m], game=game[:m], lengths=lengths, winners=winners)
        if want_pi:
            pi_y = e.empty((m, 294), torch.float32)
            if m:
                e.call("ccx_cand_to_pi", m, _p(cand), _p(pi_y))
            out["pi_y"] = pi_y
        if want_planes:
            board_x = e.empty((m, 7, 7, 7), torch.uint8)
            if m:
                e.call("ccx_encode", m, _p(state), _p(board_x), DTYPE_U8)
            out["board_x"] = board_x
        return out

    def generate_mix(self, n, normal_ratio=0.2, rand_start_ratio=0.3, **kw):
        """train_on_greedy.generate_self_play's mix (train_on_greedy.py:22-29, config.py:69-70): 20 % normal, 30 %
        random-start, the rest randomised."""
        n_norm, n_rs = int(normal_ratio * n), int(rand_start_ratio * n)
        parts = [self.generate(n_norm, **kw), self.generate(n_rs, random_start=True, **kw),
                 self.generate(n - n_norm - n_rs, randomised=True, **kw)]
        return {k: torch.cat([p[k] for p in parts], dim=-1 if k in ("state", "cand") else 0) for k in parts[0]}


class GreedyDataGenerator:
    """data_generators.py:14-80, one game per call (drop-in surface); the games are played by the batched
    generator a few hundred at a time and handed out one by one."""

    def __init__(self, randomised=False, random_start=False, engine=None, batch=256, seed=DEFAULT_SEED):
        from .engine import Engine
        self.randomised, self.random_start = randomised, random_start
        self.gen = BatchedGreedyGenerator(engine or Engine(), seed=seed)
        self.batch, self.queue = int(batch), []

    def generate_play(self):
        import numpy as np

        from .board import Board
        from .config import REWARD
        if not self.queue:
            out = self.gen.generate(self.batch, randomised=self.randomised, random_start=self.random_start, want_planes=False)
            st = out["state"].cpu().numpy().view(np.uint64)
            pi = out["pi_y"].cpu().numpy().astype(np.float64)
            lengths, winners = out["lengths"].cpu().numpy(), out["winners"].cpu().numpy()
            lo = 0
            for g in range(self.batch):
                hi = lo + int(lengths[g])
                hist = [(Board.from_packed(st[:, r]), pi[r]) for r in range(lo, hi)]
                reward = {0: REWARD['draw'], 1: REWARD['win'], 2: REWARD['lose']}[int(winners[g])]       # utils.py:34-44
                self.queue.append((hist, reward))
                lo = hi
        return self.queue.pop(0)
