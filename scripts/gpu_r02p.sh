set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_env.py -m gpu -q -x > gpurun_out/pytest_part.log 2>&1; tail -3 gpurun_out/pytest_part.log
timeout 300 python scripts/env_variants.py 65536 131072 1048576 2>&1 | tee gpurun_out/env_variants.log
