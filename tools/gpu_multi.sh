#!/bin/bash
# One multi-GPU box visit (gpurun --gpus N -- bash tools/gpu_multi.sh N): the multi-device tests, then bench.py at N ranks
# (tiles + the sample-split / in-context / one-GPU legs it carries), then the NCCL sample split as the main mode.
N=${1:-2}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi.py tests/test_gpu_sums.py -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/pytest_multi_n$N.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_n$N.json; tail -3 gpurun_out/bench_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 5 --warmup 3 --split samples --no-alt > gpurun_out/bench_n${N}_samples.json 2>> gpurun_out/bench_n$N.err
echo "bench samples rc=$?"; cut -c1-300 gpurun_out/bench_n${N}_samples.json
