set -x
mkdir -p gpurun_out
{ python scripts/selfplay_bench.py 3 4096 tc_acc; python scripts/selfplay_bench.py 3 4096 tc; python scripts/selfplay_bench.py 2 16384 tc_acc; } > gpurun_out/side_benches2.log 2>&1
cat gpurun_out/side_benches2.log
