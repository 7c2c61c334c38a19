"""Batched selfplay.selfplay() (selfplay.py:11-133): n game slots per GPU advance one ply per iteration;
opening plies are random (INITIAL_RANDOM_MOVES), later plies run a full MCTS (root pre-expanded, Dirichlet
noise, MCTS_SIMULATIONS simulations) for every slot at once with ONE batched net evaluation per simulation
round.  Finished games are labelled / discarded on the device and their slot restarts immediately.

`collect()` returns the trajectory in utils.convert_to_train_data's format: board_x (N,7,7,7), pi_y (N,294),
v_y (N,); `all_gather()` merges the per-rank buffers over NCCL (the only collective in the system)."""
import torch

from .config import (C_PUCT, DEFAULT_SEED, DIRICHLET_ALPHA, DTYPE_U8, INITIAL_RANDOM_MOVES, MCTS_SIMULATIONS,
                     PROGRESS_MOVE_LIMIT, STATE_WORDS, TOTAL_MOVES_TILL_TAU0)
from .engine import BatchedEnv, _p

NOISE_STRIDE = 128


class UniformEvaluator:
    """Stub evaluator of BASELINE configs[3]: p = 1/294 everywhere, v = 0.0."""

    def __init__(self, engine, n):
        self.p = torch.full((n, 294), 1 / 294., dtype=torch.float64, device=engine.device)
        self.v = torch.zeros((n,), dtype=torch.float64, device=engine.device)

    def __call__(self, leaf_state):
        return self.p, self.v


class BatchedSelfPlay:
    def __init__(self, engine, evaluate, n_slots=4096, seed=DEFAULT_SEED, rank=0, world=1, num_itr=MCTS_SIMULATIONS,
                 cpuct=C_PUCT, max_iters=256, edges_per_tree=0, dirichlet=True, log_moves=False, fused=None):
        self.eng, self.evaluate = engine, evaluate
        # `evaluate` = ResidualCNN.evaluate_states of a model living in this engine -> fused C round loop (ccx_mcts_run_net);
        # any other callable (stub evaluators, tests) -> one select / evaluate / expand_backup round trip per simulation
        owner = getattr(evaluate, "__self__", None)
        auto = getattr(owner, "fused_mcts", False) and getattr(owner, "eng", None) is engine and getattr(evaluate, "__name__", "") == "evaluate_states"
        self.fused = bool(auto) if fused is None else bool(fused)
        self.n, self.seed, self.rank, self.world = int(n_slots), int(seed), int(rank), int(world)
        self.num_itr, self.cpuct, self.max_iters, self.ept = int(num_itr), float(cpuct), int(max_iters), int(edges_per_tree)
        self.dirichlet = dirichlet
        e, n = engine, self.n
        self.env = BatchedEnv(n, engine=e, seed=seed, game_id0=rank * n)
        self.leaf = e.empty((5, n), torch.int64)
        self.noise = e.empty((n, NOISE_STRIDE), torch.float64)
        self.visits = e.empty((n, 294), torch.int32)
        self.tree_nodes = e.empty((n,), torch.int32)
        self.serial = e.zeros((n,), torch.int64)
        self.start_iter = e.zeros((n,), torch.int32)
        self.counters = e.zeros((8,), torch.int64)
        self.rec_state = e.zeros((self.max_iters * n, 5), torch.int64)
        self.rec_visits = e.zeros((self.max_iters * n, 294), torch.int16)
        self.rec_flag = e.zeros((self.max_iters * n,), torch.uint8)
        self.move_log = e.zeros((self.max_iters * n,), torch.int32) if log_moves else None
        self.iter = 0

    # one ply for every slot
    def step(self, restart=True):
        e, n, it = self.eng, self.n, self.iter
        if it >= self.max_iters:
            raise RuntimeError("record buffer full: collect() or raise max_iters")
        uid0 = self.rank * n
        st = self.env.state
        e.call("ccx_mcts_begin", n, _p(st), self.num_itr + 1, self.ept, INITIAL_RANDOM_MOVES)
        if self.dirichlet:
            e.call("ccx_gamma_noise", n, NOISE_STRIDE, DIRICHLET_ALPHA, self.seed, it, uid0, _p(self.noise))
        if self.fused:                                   # the library's own net: all rounds in one C call, fused round kernels
            e.call("ccx_mcts_run_net", n, self.num_itr + 1, self.cpuct, _p(self.noise) if self.dirichlet else None,
                   NOISE_STRIDE, 1)
        else:
            for r in range(self.num_itr + 1):           # round 0 = make_move's root expansion (selfplay.py:117)
                e.call("ccx_mcts_select", n, self.cpuct, _p(self.leaf))
                p, v = self.evaluate(self.leaf)
                noise = self.noise if (r == 0 and self.dirichlet) else None
                e.call("ccx_mcts_expand_backup", n, _p(p), _p(v), _p(noise), NOISE_STRIDE if noise is not None else 0, 1)
        e.call("ccx_mcts_finalize", n, 1.0, _p(self.visits), None, None, _p(self.tree_nodes))
        e.call("ccx_selfplay_advance", n, _p(st), _p(self.visits), _p(self.tree_nodes), self.seed, it, uid0, _p(self.serial),
               self.world * n, INITIAL_RANDOM_MOVES, TOTAL_MOVES_TILL_TAU0, PROGRESS_MOVE_LIMIT, _p(self.rec_state),
               _p(self.rec_visits), _p(self.rec_flag), self.max_iters, _p(self.counters), _p(self.move_log))
        e.call("ccx_selfplay_finish", n, _p(st), it, _p(self.start_iter), _p(self.serial), _p(self.rec_state),
               _p(self.rec_flag), self.max_iters, int(bool(restart)), _p(self.counters))
        self.iter += 1

    def stats(self):
        c = self.counters.cpu().tolist()
        return dict(plies=c[0], p1_wins=c[1], p2_wins=c[2], discarded_repetition=c[3], discarded_no_progress=c[4],
                    discarded_overflow=c[5], records=c[6], games=c[7], iterations=self.iter)

    def run(self, target_games=None, iters=None, poll_every=8):
        done = 0
        while True:
            self.step()
            done += 1
            if iters is not None and done >= iters:
                break
            if self.iter >= self.max_iters:
                break
            if target_games is not None and done % poll_every == 0 and self.stats()["games"] >= target_games:
                break
        return self.stats()

    # trajectory in utils.convert_to_train_data's format (utils.py:60-73)
    def collect(self):
        e = self.eng
        flags = self.rec_flag & 0xF
        rows = torch.nonzero((flags == 2) | (flags == 3)).flatten().contiguous()
        m = int(rows.numel())
        out_state = e.empty((5, max(m, 1)), torch.int64)
        pi_y = e.empty((m, 294), torch.float32)
        v_y = e.empty((m,), torch.int8)
        board_x = e.empty((m, 7, 7, 7), torch.uint8)
        if m:
            e.call("ccx_traj_pack", m, _p(rows), _p(self.rec_state), _p(self.rec_visits), _p(self.rec_flag), _p(out_state),
                   _p(pi_y), _p(v_y))
            e.call("ccx_encode", m, _p(out_state), _p(board_x), DTYPE_U8)
        return dict(board_x=board_x, pi_y=pi_y, v_y=v_y, state=out_state[:, :m])


def all_gather_trajectories(traj, group=None):
    """NCCL (or gloo) all-gather of the per-rank trajectory buffers into the training buffer; ranks hold
    different numbers of records, so counts are gathered first and buffers padded to the maximum."""
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return traj
    world = dist.get_world_size(group)
    dev = traj["board_x"].device
    m = torch.tensor([traj["board_x"].shape[0]], dtype=torch.int64, device=dev)
    counts = torch.zeros(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts, m, group=group)
    counts = counts.cpu().tolist()
    mx = max(max(counts), 1)
    out = {}
    for key in ("board_x", "pi_y", "v_y"):
        t = traj[key]
        pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
        pad[:t.shape[0]] = t
        full = torch.empty((world * mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
        dist.all_gather_into_tensor(full, pad, group=group)
        out[key] = torch.cat([full[r * mx:r * mx + counts[r]] for r in range(world)], dim=0)
    return out


# ---- selfplay.selfplay(model1, model2=None, randomised=False), one game, same contract as selfplay.py:11-80 ----
def selfplay(model1, model2=None, randomised=False, num_itr=MCTS_SIMULATIONS):
    """Returns (play_history [(Board, pi)], reward for player 1) or (None, None) for a discarded game; what
    train.py:62 calls.  For throughput use BatchedSelfPlay; this mirror exists for drop-in compatibility."""
    import copy
    import random

    import numpy as np

    from . import utils
    from .board import Board
    from .config import (BOARD_HIST_MOVES, DET_TREE_TAU, DIR_NOISE_FACTOR, NUM_CHECKERS, PLAYER_ONE, PLAYER_TWO,
                         TOTAL_HIST_MOVES, TREE_TAU, UNIQUE_DEST_LIMIT)
    from .MCTS import MCTS, Node
    model2 = model2 or model1
    player_progresses, player_turn, num_useless_moves, play_history, tree_tau = [0, 0], 0, 0, [], TREE_TAU
    root = Node(Board(randomised=randomised), PLAYER_ONE)
    use_model1 = True
    while True:
        model = model1 if use_model1 else model2
        if len(root.state.hist_moves) < INITIAL_RANDOM_MOVES:                       # selfplay.py:32-33, 83-104
            valid = root.state.get_valid_moves(root.currPlayer)
            start = random.choice([k for k in valid if valid[k]])
            nxt = copy.deepcopy(root.state)
            nxt.place(root.currPlayer, start, random.choice(valid[start]))
            root = Node(nxt, PLAYER_ONE + PLAYER_TWO - root.currPlayer)
        else:                                                                       # selfplay.py:107-133
            tree = MCTS(root, model, num_itr=num_itr, tree_tau=tree_tau)
            tree.expandAndBackUp(tree.root, breadcrumbs=[])
            noise = np.random.dirichlet(np.ones(len(tree.root.edges)) * DIRICHLET_ALPHA)
            for i, e in enumerate(tree.root.edges):
                e.stats['P'] *= (1. - DIR_NOISE_FACTOR)
                e.stats['P'] += DIR_NOISE_FACTOR * noise[i]
            pi, edge = tree.search()
            play_history.append((tree.root.state, pi))
            root = Node(copy.deepcopy(edge.outNode.state), edge.outNode.currPlayer)
        hm = root.state.hist_moves
        mine = [hm[i] for i in range(len(hm) - 1, -1, -2)]
        if len(mine) * 2 >= TOTAL_HIST_MOVES and len(set(m[1] for m in mine)) <= UNIQUE_DEST_LIMIT:
            return None, None                                                       # selfplay.py:45-47
        progress = root.state.player_progress(player_turn + 1)
        if progress > player_progresses[player_turn]:
            num_useless_moves = int(num_useless_moves * (NUM_CHECKERS - 1) / NUM_CHECKERS)
            player_progresses[player_turn] = progress
        else:
            num_useless_moves += 1
        player_turn, use_model1 = 1 - player_turn, not use_model1
        if len(play_history) + INITIAL_RANDOM_MOVES > TOTAL_MOVES_TILL_TAU0:
            tree_tau = DET_TREE_TAU
        if root.state.check_win():
            break
        if num_useless_moves >= PROGRESS_MOVE_LIMIT:
            return None, None
    reward = utils.get_p1_winloss_reward(root.state)
    return (play_history[BOARD_HIST_MOVES:] if randomised else play_history), reward
