"""cfg 5 across ranks on real GPUs (BASELINE configs[4]): every rank plays its share of self-play games to completion in the
accurate net mode, the trajectories are all-gathered over NCCL with their real, ragged counts (train.py:88-92) and every
rank checks the gathered buffer.  Needs >= 2 CUDA devices (`gpurun --gpus 2`); skipped on a 1-GPU box."""
import os
import sys

import pytest
import torch

from conftest import GOLDEN, ROOT

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    sys.path.insert(0, ROOT)
    from chinesecheckersagent_b200.engine import Engine
    from chinesecheckersagent_b200.model import ResidualCNN
    from chinesecheckersagent_b200.selfplay import BatchedSelfPlay, all_gather_trajectories, check_gathered_trajectories
    eng = Engine(rank)
    model = ResidualCNN(engine=eng).load_weights(os.path.join(GOLDEN, "good_model_weights.npz"))
    assert model.kernel == "tc_acc"
    games = 300 + 100 * rank                                   # uneven shares -> ragged record counts
    sp = BatchedSelfPlay(eng, model.evaluate_states, n_slots=128, seed=11, rank=rank, world=world, num_itr=24, max_iters=400)
    st = sp.play_games(games)
    local = sp.collect()
    gathered = all_gather_trajectories(local)
    chk = check_gathered_trajectories(eng, local, gathered)
    q.put((rank, st, chk, int(local["v_y"].shape[0])))
    dist.barrier()
    dist.destroy_process_group()


def test_selfplay_to_completion_then_nccl_all_gather_checked_on_every_rank():
    world = torch.cuda.device_count()
    if world < 2:
        pytest.skip("needs >= 2 CUDA devices")
    world = 2
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=600) for _ in range(world)), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    totals = {r[2]["records_total"] for r in res}
    assert len(totals) == 1 and totals.pop() == sum(r[3] for r in res) > 0
    assert res[0][2]["checksum"] == res[1][2]["checksum"]
    assert res[0][2]["counts"] == [res[0][3], res[1][3]] and res[0][3] != res[1][3]
    for rank, st, chk, _ in res:
        assert st["unfinished"] == 0 and st["games_started"] == 300 + 100 * rank
        ended = st["p1_wins"] + st["p2_wins"] + st["discarded_repetition"] + st["discarded_no_progress"] + st["discarded_overflow"]
        assert ended == st["games_started"] and st["games"] > 0.5 * st["games_started"]
        assert chk["legal_support_checked"] > 0
