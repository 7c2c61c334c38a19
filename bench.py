#!/usr/bin/env python
"""bench.py — headline benchmark of the ccx engine (BASELINE.json metric: env steps/sec).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ccx|reference]

A "step" is one pass of the fused env kernel over one batch: 65,536 games/GPU x 256 random-legal plies
(BASELINE configs[1], SURVEY.md §8d cfg 2).  See DESIGN.md §Measurement for every field of the JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GAMES_PER_GPU = 65536
PLIES_PER_STEP = 256
BYTES_PER_ENV_STEP = 80          # 40 B state read + 40 B written per game-ply
SEED = 0x5EED2026
METRIC = "env_steps_per_sec"
UNIT = "env steps/s"


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def measured_peak_tflops():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["bf16_tflops"])
    except Exception:
        return 1590.0


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], stdout=subprocess.PIPE, text=True, timeout=5).stdout
                self.samples.append([x.strip() for x in out.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        self.stop_flag = True
        sm, mx, reasons = [], 0.0, set()
        for s in self.samples:
            try:
                sm.append(float(s[0])); mx = max(mx, float(s[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_baseline(nthreads, seconds_target=12.0):
    """Oracle port (C restatement of board.py) timed on the host cores on a bounded sample."""
    import oracle as orc
    games, plies = 4096, 32
    st = orc.start_states(games)
    t = time.perf_counter(); orc.step_random(st, SEED, 0, plies, nthreads=nthreads); dt = time.perf_counter() - t
    rate = games * plies / dt
    games = int(min(GAMES_PER_GPU, max(4096, rate * seconds_target / PLIES_PER_STEP)))
    st = orc.start_states(games)
    t = time.perf_counter(); orc.step_random(st, SEED, 0, PLIES_PER_STEP, nthreads=nthreads); dt = time.perf_counter() - t
    return games * PLIES_PER_STEP / dt, "%d games x %d plies from the start position, %d thread(s)" % (games, PLIES_PER_STEP, nthreads)


def run_reference(args, rank):
    """--impl reference: the reference's CPU path for this workload on the box's host cores.  The
    reference is Python and cannot travel to the GPU box, so this times the oracle port
    (oracle/ccx_oracle.c, pinned to the reference by tests/golden) with every host thread."""
    if rank != 0:
        return
    import oracle as orc
    cores = os.cpu_count() or 1
    games = 8192
    st = orc.start_states(games)
    for _ in range(args.warmup):
        st, _, _ = orc.step_random(st, SEED, 0, 16, nthreads=cores)
    t = time.perf_counter()
    for k in range(args.steps):
        st, _, _ = orc.step_random(st, SEED, k * PLIES_PER_STEP, PLIES_PER_STEP, nthreads=cores)
    dt = time.perf_counter() - t
    value = games * PLIES_PER_STEP * args.steps / dt
    sample = "%d games x %d plies per step (bounded sample of the 65,536-game workload)" % (games, PLIES_PER_STEP)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": "cfg2 random-legal env stepping", "games_per_gpu": GAMES_PER_GPU, "plies_per_step": PLIES_PER_STEP},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ccx", choices=["ccx", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--games", type=int, default=GAMES_PER_GPU, help="games per GPU (BASELINE config: 65536)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ccx" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from chinesecheckersagent_b200.engine import BatchedEnv, Engine, HostEnv

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    eng = Engine(local_rank)
    n = args.games
    env = BatchedEnv(n, engine=eng, seed=SEED, game_id0=rank * n)     # global game ids: sharding-invariant RNG
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        env.step_random(PLIES_PER_STEP)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = eng.launches
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for a, b in evs:
        flush.fill_(1)                      # L2 flush between timed iterations (not timed)
        a.record()
        env.step_random(PLIES_PER_STEP)     # torch's current stream == the handle's stream
        b.record()
    barrier()
    launches = eng.launches - launches0
    per_step_ms = [a.elapsed_time(b) for a, b in evs]
    total_ms = torch.tensor([sum(per_step_ms)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_s = float(total_ms.item()) / 1e3
    clocks = sampler.summary() if rank == 0 else None
    steps_total = world * n * PLIES_PER_STEP * args.steps
    value = steps_total / total_s

    # ---- end-to-end through the host-buffer C-ABI (pinned host state in, state + win counters out)
    host = HostEnv(engine=eng)
    st_pinned = torch.from_numpy(np.ascontiguousarray(env.numpy_state()).view(np.int64)).pin_memory()
    st_np = st_pinned.numpy().view(np.uint64)
    wins = np.zeros(2, dtype=np.uint64)
    e2e_steps = max(3, min(args.steps, 10))
    host.step_random(st_np, PLIES_PER_STEP, seed=SEED, step0=0, game_id0=rank * n, wins=wins)
    barrier()
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        host.step_random(st_np, PLIES_PER_STEP, seed=SEED, step0=(k + 1) * PLIES_PER_STEP, game_id0=rank * n, wins=wins)
    torch.cuda.synchronize()
    e2e_t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = world * n * PLIES_PER_STEP * e2e_steps / float(e2e_t.item())
    h2d = n * 40 + 16
    d2h = n * 40 + 16

    extra = {}
    if not args.no_extra:
        try:
            from chinesecheckersagent_b200 import bench_extra
            extra = bench_extra.run(eng, rank, world, barrier, peak_gbs=measured_peak_gbs()[0], peak_tflops=measured_peak_tflops())
        except ImportError:
            pass

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        launch_ms = sum(per_step_ms) / len(per_step_ms)
        achieved = BYTES_PER_ENV_STEP * n * PLIES_PER_STEP / (launch_ms * 1e-3) / 1e9
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                traffic = json.load(f).get("k_step_random_flat")
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_s / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic",
            "config": {"workload": "cfg2 random-legal env stepping (movegen+pick+apply+win)", "games_per_gpu": n,
                       "plies_per_step": PLIES_PER_STEP, "start": "Board() start position, won games restart",
                       "l2": "flushed between timed steps (256 MiB write, untimed)", "parallelism": "games sharded, no collective"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel": "k_step_random_flat<false>",
                         "algorithmic_bytes_per_launch": BYTES_PER_ENV_STEP * n * PLIES_PER_STEP,
                         "launch_ms": launch_ms},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "ccx_step_random_host (pinned host state in/out)", "steps": e2e_steps},
            "gpu_launches": launches, "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            v1, sample = cpu_baseline(1)
            line["cpu_baseline"] = {"value": v1, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample,
                                    "note": "C restatement of board.py (oracle/); the Python reference itself measured ~4.2e3 steps/s/core (BASELINE.md)"}
        if extra:
            line["extra"] = extra
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
