#!/bin/bash
# gpurun --gpus N payload: multi-GPU tests + bench at N ranks (torchrun) + per-rank stage table
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_n$N.txt
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_multirank_nccl.py -x -q > gpurun_out/pytest_n$N.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_n$N.log
tail -4 gpurun_out/pytest_n$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo "bench rc=$?"
tail -c 300 gpurun_out/r02_bench_n$N.err
head -c 1500 gpurun_out/r02_bench_n$N.json
