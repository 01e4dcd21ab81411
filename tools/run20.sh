#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_batched_affine.py tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/pytest20.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest20.log
tail -5 gpurun_out/pytest20.log
timeout 900 python tools/accumulate_probe.py 21,22,23,24 0 "512,16,16,48;512,16,8,48;512,16,4,48;512,48,16,48;1024,16,16,48;256,16,16,48" > gpurun_out/ba_v2_20.txt 2>&1
grep -E "^---|mode=|equal" gpurun_out/ba_v2_20.txt | cut -c1-100
