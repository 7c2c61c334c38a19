set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_env.py tests/test_bench_contract.py -m gpu -q -x > gpurun_out/pytest_part.log 2>&1; tail -5 gpurun_out/pytest_part.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
