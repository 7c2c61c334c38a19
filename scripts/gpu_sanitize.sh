# compute-sanitizer over every kernel family (scripts/sanitize.py): memcheck, racecheck, synccheck
set -x
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool python scripts/sanitize.py > gpurun_out/sanitize_$tool.log 2>&1
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize pass done" gpurun_out/sanitize_$tool.log | tail -2
done
