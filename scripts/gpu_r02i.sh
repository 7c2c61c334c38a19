set -x
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 1200 compute-sanitizer --tool $tool python scripts/sanitize.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize pass done" gpurun_out/sanitize_$tool.log | tail -3
done
