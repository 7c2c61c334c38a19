set -x
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck python scripts/sanitize.py > gpurun_out/sanitize_memcheck.log 2>&1; grep -E "ERROR SUMMARY|sanitize pass done" gpurun_out/sanitize_memcheck.log | tail -2
timeout 900 compute-sanitizer --tool racecheck python scripts/sanitize.py > gpurun_out/sanitize_racecheck.log 2>&1; grep -E "RACECHECK SUMMARY|sanitize pass done" gpurun_out/sanitize_racecheck.log | tail -2
timeout 900 compute-sanitizer --tool synccheck python scripts/sanitize.py > gpurun_out/sanitize_synccheck.log 2>&1; grep -E "ERROR SUMMARY|sanitize pass done" gpurun_out/sanitize_synccheck.log | tail -2
