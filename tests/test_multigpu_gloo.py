"""Host-side logic of the N>1 path on CPU with the gloo backend, world_size 2: trajectory all-gather with
ragged per-rank counts, global game ids per rank (sharding-invariant RNG keys), and bench.py's reference arm
running on rank 0 only."""
import json
import os
import subprocess
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from chinesecheckersagent_b200.selfplay import all_gather_trajectories
    m = 3 + 4 * rank                                         # ragged: 3 records on rank 0, 7 on rank 1
    traj = dict(board_x=torch.full((m, 7, 7, 7), rank + 1, dtype=torch.uint8),
                pi_y=torch.full((m, 294), float(rank + 1)), v_y=torch.full((m,), 1 - 2 * rank, dtype=torch.int8))
    out = all_gather_trajectories(traj)
    q.put((rank, out["board_x"].shape[0], out["board_x"][:, 0, 0, 0].tolist(), out["pi_y"][:, 0].tolist(), out["v_y"].tolist()))
    dist.destroy_process_group()


def test_trajectory_all_gather_ragged_counts():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, n, bx, pi, vy in res:
        assert n == 10
        assert bx == [1] * 3 + [2] * 7 and pi == [1.0] * 3 + [2.0] * 7 and vy == [1] * 3 + [-1] * 7


def _check_worker(rank, world, port, q, tamper):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from chinesecheckersagent_b200.selfplay import all_gather_trajectories, check_gathered_trajectories
    g = torch.Generator().manual_seed(100 + rank)
    m = 5 + 6 * rank                                          # ragged
    pi = torch.rand((m, 294), generator=g)
    traj = dict(board_x=torch.randint(0, 7, (m, 7, 7, 7), generator=g, dtype=torch.uint8), pi_y=pi / pi.sum(1, keepdim=True),
                v_y=(torch.randint(0, 2, (m,), generator=g) * 2 - 1).to(torch.int8))
    out = all_gather_trajectories(traj)
    if tamper and rank == 1:
        out["board_x"][0, 0, 0, 0] += 1                       # rank 1 now holds a different buffer than rank 0
    try:
        res = check_gathered_trajectories(None, traj, out)    # no `state` key: the legal-move check (GPU movegen) is skipped
        q.put((rank, "ok", res["records_total"], res["counts"]))
    except AssertionError as e:
        q.put((rank, "fail", str(e)[:60], None))
    dist.destroy_process_group()


def _run_check(tamper):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000 + (7 if tamper else 0)
    procs = [ctx.Process(target=_check_worker, args=(r, 2, port, q, tamper)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    return res


def test_gathered_buffer_checks_pass_and_catch_a_diverging_rank():
    """check_gathered_trajectories (what bench.py's cfg 5 leg and the 2-GPU test assert on every rank): count = sum of the ranks'
    counts, own slice intact, same checksum everywhere; a rank whose buffer differs is detected on both ranks."""
    ok = _run_check(False)
    assert [r[1] for r in ok] == ["ok", "ok"] and ok[0][2] == 16 and ok[0][3] == [5, 11]
    bad = _run_check(True)
    assert [r[1] for r in bad] == ["fail", "fail"]


def test_single_process_gather_is_identity():
    sys.path.insert(0, ROOT)
    from chinesecheckersagent_b200.selfplay import all_gather_trajectories
    t = dict(board_x=torch.zeros((2, 7, 7, 7), dtype=torch.uint8), pi_y=torch.zeros((2, 294)), v_y=torch.zeros(2, dtype=torch.int8))
    assert all_gather_trajectories(t) is t


def test_reference_arm_runs_on_rank0_only():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "0", "--no-extra"], env=env, stdout=subprocess.PIPE, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip() == ""
    env["RANK"] = "0"
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "0", "--no-extra"], env=env, stdout=subprocess.PIPE, text=True, timeout=300)
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "env_steps_per_sec" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["e2e"]["h2d_bytes_per_step"] == 0


def test_game_ids_are_global_across_ranks():
    """Rank r owns game ids [r*n, (r+1)*n): the oracle (same Philox keys as the kernels) gives identical
    results for a shard computed alone and as part of the whole."""
    import oracle as orc
    whole, _, _ = orc.step_random(orc.start_states(64), 11, 0, 20, game_id0=0)
    shard, _, _ = orc.step_random(orc.start_states(32), 11, 0, 20, game_id0=32)
    assert np.array_equal(whole[:, 32:], shard)
