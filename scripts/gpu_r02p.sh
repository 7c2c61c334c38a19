set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_net.py -m gpu -q -x > gpurun_out/pytest_part.log 2>&1; tail -3 gpurun_out/pytest_part.log
python scripts/dropin_bench.py; CCX_ACC_CTX=3 python scripts/dropin_bench.py; CCX_ACC_CTX=0 python scripts/dropin_bench.py
for n in 1; do python scripts/net_bench.py 20 tc_acc; done
python scripts/fullplay_bench.py | tail -2
