set -x
mkdir -p gpurun_out
python scripts/greedy_variants.py > gpurun_out/greedy_variants.log 2>&1; cat gpurun_out/greedy_variants.log
