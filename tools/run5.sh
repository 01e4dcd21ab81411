#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pasta.py -x -q > gpurun_out/pytest5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest5.log
tail -25 gpurun_out/pytest5.log
timeout 600 python tools/pasta_probe.py 16,20,22,24 > gpurun_out/pasta_probe.txt 2>&1
tail -12 gpurun_out/pasta_probe.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_batched_affine.py -x -q > gpurun_out/pytest5b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest5b.log
tail -3 gpurun_out/pytest5b.log
