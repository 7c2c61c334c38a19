set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_selfplay.py tests/test_gpu_dropin_reference.py tests/test_gpu_mcts.py tests/test_gpu_arena.py -m gpu -q -x > gpurun_out/pytest_part.log 2>&1; tail -15 gpurun_out/pytest_part.log
