set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_env.py -m gpu -q -x > gpurun_out/pytest_part.log 2>&1; tail -3 gpurun_out/pytest_part.log
timeout 900 compute-sanitizer --tool memcheck python scripts/sanitize.py > gpurun_out/sanitize_memcheck.log 2>&1; grep -E "ERROR SUMMARY|sanitize pass done" gpurun_out/sanitize_memcheck.log | tail -2
timeout 900 compute-sanitizer --tool racecheck python scripts/sanitize.py > gpurun_out/sanitize_racecheck.log 2>&1; grep -E "RACECHECK SUMMARY|sanitize pass done" gpurun_out/sanitize_racecheck.log | tail -2
