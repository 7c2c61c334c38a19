set -x
mkdir -p gpurun_out
timeout 600 python scripts/env_variants.py 65536 131072 1048576 > gpurun_out/env_variants2.log 2>&1; cat gpurun_out/env_variants2.log
