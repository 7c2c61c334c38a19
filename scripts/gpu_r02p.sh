set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_net.py -m gpu -q -x -k "split" > gpurun_out/pytest_part.log 2>&1; tail -3 gpurun_out/pytest_part.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 200 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
