#!/bin/bash
# One GPU pass: parity tests, bench line, ncu launch list + full captures of the dominant kernels.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_env.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/ncu_env.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_full.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_step_random -s 3 -c 1 -o gpurun_out/prof_step_random python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/ncu_p1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_net_trunk|k_policy_dense|k_mcts_search' -c 4 -o gpurun_out/prof_net_mcts python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_p2.log 2>&1
ls -la gpurun_out
