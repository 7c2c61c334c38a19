set -x
mkdir -p gpurun_out
export CCX_LIB_PATH=$PWD/chinesecheckersagent_b200/libccx_timing.so
( python scripts/acc_timing.py 1 4; python scripts/acc_timing.py 1 65536; python scripts/acc_timing.py 3 65536 ) > gpurun_out/acc_timing.log 2>&1
tail -3 gpurun_out/acc_timing.log | cut -c1-300
