#!/bin/bash
# gpurun payload: chain-kernel tests + probes; everything lands in gpurun_out/
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_batched_affine.py tests/test_gpu_sort_path.py tests/test_gpu_external_kat.py -x -q > gpurun_out/pytest2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest2.log
tail -8 gpurun_out/pytest2.log
timeout 600 python tools/chain_probe.py 20,21,22,24 0 8,12,16 > gpurun_out/chain_probe.txt 2>&1
tail -30 gpurun_out/chain_probe.txt
