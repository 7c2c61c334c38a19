set -x
mkdir -p gpurun_out
timeout 600 python scripts/env_variants.py 65536 > gpurun_out/env_variants.log 2>&1; cat gpurun_out/env_variants.log
