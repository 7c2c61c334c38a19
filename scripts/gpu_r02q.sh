set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_net_trunk_acc" -s 3 -c 1 -o gpurun_out/prof_net_accm python scripts/net_bench.py 1 tc_acc > gpurun_out/ncu_p5.log 2>&1
{ python scripts/selfplay_bench.py 3 4096 tc_acc; python scripts/selfplay_bench.py 3 16384 tc_acc; python scripts/net_bench.py 10 tc_acc; python scripts/fullplay_bench.py; } > gpurun_out/side_benches.log 2>&1
cat gpurun_out/side_benches.log
