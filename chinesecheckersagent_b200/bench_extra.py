"""Secondary measurements reported under "extra" in bench.py's JSON line: the other BASELINE configs
(cfg 3 greedy self-play, cfg 4 stub MCTS, plane encoder).  Device-timed with CUDA events on the
launching stream, max over ranks."""
import json
import os

import torch

from .config import DEFAULT_SEED, DTYPE_BF16
from .engine import BatchedEnv, BatchedMCTS

MCTS_TREES = 4096            # BASELINE configs[3]
MCTS_SIMS = 175              # config.py:35
MCTS_BYTES_PER_SIM = 1400    # ~0.9 KB read + 0.5 KB written per simulation
GREEDY_GAMES = 131072        # BASELINE configs[2]: 1M games / 8 GPUs


def _timed(fn, reps, world):
    evs = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    t = torch.tensor([sum(a.elapsed_time(b) for a, b in evs) / 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def run(eng, rank, world, barrier, peak_gbs=None, peak_tflops=None):
    out = {}
    # ---- cfg 4: MCTS, uniform-prior stub evaluator, 4,096 concurrent trees per GPU --------------------
    env = BatchedEnv(MCTS_TREES, engine=eng, seed=DEFAULT_SEED, game_id0=rank * MCTS_TREES)
    env.step_random(6)                                   # roots = start advanced by 6 random plies
    mcts = BatchedMCTS(eng, num_itr=MCTS_SIMS)
    for _ in range(3):                                   # warm-up: pool allocation, and torch's caching allocator gets the
        res = mcts.search(env.state)                     # output blocks a steady-state caller reuses (no cudaMalloc while timing)
    stats = dict(overflowed=int((res["n_nodes"] < 0).sum().item()), mean_nodes=float(res["n_nodes"].float().mean().item()))
    del res
    barrier()
    l0 = eng.launches
    reps = 10
    t = _timed(lambda: mcts.search(env.state), reps, world)
    sims = world * MCTS_TREES * MCTS_SIMS * reps
    out["mcts_stub"] = {"metric": "mcts_sims_per_sec", "value": sims / t, "unit": "sims/s", "trees_per_gpu": MCTS_TREES,
                        "sims_per_move": MCTS_SIMS, "ms_per_search": t / reps * 1e3, "evaluator": "uniform prior 1/294, v=0",
                        "gpu_launches": eng.launches - l0,
                        "overflowed_trees": stats["overflowed"], "mean_nodes_per_tree": stats["mean_nodes"]}
    if peak_gbs:
        ach = MCTS_BYTES_PER_SIM * MCTS_TREES * MCTS_SIMS / (t / reps) / 1e9
        out["mcts_stub"]["roofline"] = {"bound": "hbm", "achieved": ach, "peak": peak_gbs, "unit": "GB/s", "frac": ach / peak_gbs,
                                        "kernel": "k_mcts_search<uniform>"}
    # ---- cfg 3: greedy-vs-greedy self-play to the end (game.py:58-100) ----------------------------------
    genv = BatchedEnv(GREEDY_GAMES, engine=eng, seed=DEFAULT_SEED, game_id0=rank * GREEDY_GAMES)
    genv.play_greedy()
    barrier()
    counters = eng.zeros((4,), torch.int64)

    def greedy_once():
        genv.reset()
        genv.play_greedy(counters=counters)
    reps = 3
    t = _timed(greedy_once, reps, world)
    c = counters.cpu().tolist()
    out["greedy_selfplay"] = {"metric": "env_steps_per_sec", "value": world * c[0] / t, "unit": "plies/s",
                              "games_per_sec": world * GREEDY_GAMES * reps / t, "games_per_gpu": GREEDY_GAMES,
                              "mean_plies_per_game": c[0] / (GREEDY_GAMES * reps), "p1_win_frac": c[1] / (GREEDY_GAMES * reps),
                              "p2_win_frac": c[2] / (GREEDY_GAMES * reps), "repetition_stop_frac": c[3] / (GREEDY_GAMES * reps)}
    # ---- §8f f3: greedy supervised-data generator (data_generators.py), records + pi + planes on the device ---------
    from .data_generators import BatchedGreedyGenerator
    gen = BatchedGreedyGenerator(eng, seed=DEFAULT_SEED, rank=rank, world=world)
    n_rec = 0
    DG = GREEDY_GAMES // 4                               # 32,768 games per call: ~1.4 M records, 1.7 GB of pi_y per call
    for _ in range(3):                                   # warm-up (the caching allocator ends up holding the output blocks)
        n_rec = int(gen.generate(DG)["v_y"].shape[0])
    barrier()
    t = _timed(lambda: gen.generate(DG), 4, world)
    out["greedy_datagen"] = {"metric": "games_per_sec", "value": world * DG * 4 / t, "unit": "games/s",
                             "records_per_sec_approx": world * n_rec * 4 / t, "games_per_gpu": DG,
                             "outputs": "board_x u8 (M,7,7,7), pi_y f32 (M,294), v_y i8 (M,)"}
    del gen
    # ---- plane encoder (utils.to_model_input) straight into a bf16 NHWC tensor ----------------------------
    eenv = BatchedEnv(1 << 20, engine=eng, seed=DEFAULT_SEED)
    eenv.step_random(8)
    planes = eng.empty((1 << 20, 7, 7, 7), torch.bfloat16)
    eenv.encode(DTYPE_BF16, out=planes)
    barrier()
    reps = 10
    t = _timed(lambda: eenv.encode(DTYPE_BF16, out=planes), reps, world)
    n = (1 << 20) * reps * world
    out["encode_bf16"] = {"metric": "positions_per_sec", "value": n / t, "unit": "positions/s",
                          "bytes_per_position": 40 + 686, "achieved_gbs": (40 + 686) * (1 << 20) * reps / t / 1e9}
    if peak_gbs:
        out["encode_bf16"]["frac_of_hbm_peak"] = out["encode_bf16"]["achieved_gbs"] / peak_gbs
    # ---- cfg 5: full self-play, MCTS + good_model policy/value net, 4,096 concurrent trees per GPU ---------------
    weights = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "good_model_weights.npz")
    if os.path.exists(weights):
        from .model import ResidualCNN
        from .selfplay import BatchedSelfPlay, all_gather_trajectories
        model = ResidualCNN(engine=eng).load_weights(weights).set_kernel("tc")       # the throughput mode; the accurate default is timed below
        sp = BatchedSelfPlay(eng, model.evaluate_states, n_slots=MCTS_TREES, seed=DEFAULT_SEED, rank=rank, world=world,
                             num_itr=MCTS_SIMS, max_iters=16)
        for _ in range(7):                                # 6 opening plies + 1 searched ply as warm-up
            sp.step()
        barrier()
        l0 = eng.launches
        reps = 2
        t = _timed(sp.step, reps, world)
        sims = world * MCTS_TREES * MCTS_SIMS * reps
        evals = world * MCTS_TREES * (MCTS_SIMS + 1) * reps
        out["selfplay_net"] = {"metric": "mcts_sims_per_sec", "value": sims / t, "unit": "sims/s", "net_evals_per_sec": evals / t,
                               "trees_per_gpu": MCTS_TREES, "sims_per_move": MCTS_SIMS, "ms_per_ply_iteration": t / reps * 1e3,
                               "net": "good_model.h5 (fixture copy), tcgen05 kernels, fp16 operands / fp32 accumulate",
                               "gpu_launches": eng.launches - l0}
        # the same self-play step with four times the slots: the net runs at its large-batch rate and the tree kernels have
        # 4x the warps to hide their latency chains behind (BASELINE names 4,096 trees/GPU; this is the throughput setting)
        del sp
        big = BatchedSelfPlay(eng, model.evaluate_states, n_slots=4 * MCTS_TREES, seed=DEFAULT_SEED, rank=rank, world=world,
                              num_itr=MCTS_SIMS, max_iters=12)
        for _ in range(7):
            big.step()
        barrier()
        t = _timed(big.step, 2, world)
        out["selfplay_net_16k_slots"] = {"metric": "mcts_sims_per_sec", "value": world * 4 * MCTS_TREES * MCTS_SIMS * 2 / t, "unit": "sims/s",
                                         "trees_per_gpu": 4 * MCTS_TREES, "ms_per_ply_iteration": t / 2 * 1e3}
        traj_src = big
        # the reference-facing drop-in surface, one game at a time: AiPlayer.decide_move's MCTS(Node(Board()), model).search()
        # (player.py:157-158) through the Python mirror -- what a user gets by only swapping the imports
        from .board import Board
        from .MCTS import MCTS, Node
        import time
        def one_decision():
            return MCTS(Node(Board(engine=eng), 1), model, num_itr=MCTS_SIMS).search()
        one_decision()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            one_decision()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 3
        out["dropin_single_game_search"] = {"metric": "ms_per_move_decision", "value": dt * 1e3, "unit": "ms", "sims_per_sec": MCTS_SIMS / dt,
                                            "api": "MCTS(Node(Board(), 1), ResidualCNN).search(), batch of one tree",
                                            "reference_cpu": "~63 sims/s/core with a torch-CPU restatement of the net (BASELINE.md §2)"}
        # net forward alone on a big batch
        planes = torch.randint(0, 7, (65536, 7, 7, 7), dtype=torch.uint8, device=eng.device)
        for kern in ("tc", "tc_acc", "simt"):
            model.set_kernel(kern)
            model.forward(planes)
            barrier()
            reps = 5
            t = _timed(lambda: model.forward(planes), reps, world)
            tf = 6.483264e6 * 65536 * reps / t / 1e12 * world
            out["net_forward_" + kern] = {"metric": "positions_per_sec", "value": world * 65536 * reps / t, "unit": "positions/s",
                                          "tflops": tf, "batch": 65536}
            if kern == "tc" and peak_tflops:
                out["net_forward_tc"]["roofline"] = {"bound": "tensor", "achieved": tf / world, "peak": peak_tflops, "unit": "TFLOP/s",
                                                     "frac": tf / world / peak_tflops, "kernel": "k_net_trunk_tc4 + k_policy_dense_tc3"}
        # cfg 5 with the accurate (split-precision) net: the mode in which the <= 1e-3 net bar and the >= 1e7 sims/s bar hold together
        model.set_kernel("tc_acc")
        acc_sp = BatchedSelfPlay(eng, model.evaluate_states, n_slots=MCTS_TREES, seed=DEFAULT_SEED, rank=rank, world=world,
                                 num_itr=MCTS_SIMS, max_iters=12)
        for _ in range(7):
            acc_sp.step()
        barrier()
        t = _timed(acc_sp.step, 2, world)
        out["selfplay_net_accurate"] = {"metric": "mcts_sims_per_sec", "value": world * MCTS_TREES * MCTS_SIMS * 2 / t, "unit": "sims/s",
                                        "trees_per_gpu": MCTS_TREES, "ms_per_ply_iteration": t / 2 * 1e3,
                                        "net": "tc_acc: split-precision tcgen05 kernels (max |dp| 2e-5 vs the float64 restatement)"}
        del acc_sp
        model.set_kernel("tc")
        # trajectory all-gather (the only collective): time it when there is more than one rank
        traj = traj_src.collect()
        synthetic = int(traj["board_x"].shape[0]) == 0
        if synthetic:        # no game finishes within the few plies timed above: gather a buffer the size of ~16 plies of 4,096 slots
            m = 65536
            traj = dict(board_x=torch.zeros((m, 7, 7, 7), dtype=torch.uint8, device=eng.device),
                        pi_y=torch.zeros((m, 294), dtype=torch.float32, device=eng.device),
                        v_y=torch.zeros((m,), dtype=torch.int8, device=eng.device))
        if world > 1:
            all_gather_trajectories(traj)
            barrier()
            t = _timed(lambda: all_gather_trajectories(traj), 3, world)
            rec = int(traj["board_x"].shape[0])
            out["trajectory_all_gather"] = {"ms": t / 3 * 1e3, "records_this_rank": rec, "synthetic_buffer": synthetic,
                                            "bytes_per_rank": rec * (343 + 294 * 4 + 1), "backend": "nccl"}
    return out
