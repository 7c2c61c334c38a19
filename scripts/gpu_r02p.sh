set -x
mkdir -p gpurun_out
timeout 300 python scripts/env_variants.py 65536 2>&1 | tee gpurun_out/env_variants.log
