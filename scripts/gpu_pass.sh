#!/bin/bash
# One GPU pass: parity tests, smoke, bench lines (both arms), ncu launch lists + full captures of the dominant kernels.
# Afterwards, on the CPU box: python scripts/traffic_from_ncu.py <tag>   (writes profiles/traffic.json from the captures)
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 300 gpurun_out/bench_ref.json
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 300 gpurun_out/bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_env.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/ncu_env.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3900 -c 1100 --csv --log-file gpurun_out/launches_sp.csv python scripts/selfplay_bench.py 1 4096 tc_acc > gpurun_out/ncu_sp.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_step_random -s 3 -c 1 -o gpurun_out/prof_step_random python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/ncu_p1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_mcts_search -c 1 -o gpurun_out/prof_mcts_search python scripts/mcts_stub_bench.py > gpurun_out/ncu_p2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_net_trunk_acc|k_policy_dense_acc" -s 6 -c 2 -o gpurun_out/prof_net_acc python scripts/net_bench.py 1 tc_acc > gpurun_out/ncu_p5.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_mcts_round" -s 1320 -c 1 -o gpurun_out/prof_tree python scripts/selfplay_bench.py 1 4096 tc_acc > gpurun_out/ncu_p4.log 2>&1
{ python scripts/selfplay_bench.py 3 4096 tc_acc; python scripts/selfplay_bench.py 3 4096 tc; python scripts/net_bench.py 10 tc; python scripts/net_bench.py 10 tc_acc; python scripts/mcts_stub_bench.py; } > gpurun_out/side_benches.log 2>&1
cat gpurun_out/side_benches.log
ls -la gpurun_out
