# multi-GPU bench line: bash scripts/gpu_multi.sh N   (run under gpurun --gpus N)
set -x
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_gpu_multigpu.py -m gpu -q -x > gpurun_out/pytest_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_2gpu.log
  tail -3 gpurun_out/pytest_2gpu.log
fi
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; tail -c 600 gpurun_out/bench_${N}gpu.json; tail -5 gpurun_out/bench_${N}gpu.err
