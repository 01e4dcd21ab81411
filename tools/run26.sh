#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_poseidon.py tests/test_plonk_verifier.py -x -q -m gpu > gpurun_out/pytest26.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest26.log
tail -4 gpurun_out/pytest26.log
timeout 300 python tools/plonk_batch_stages.py > gpurun_out/plonk_batch_stages26.txt 2>&1; cut -c1-200 gpurun_out/plonk_batch_stages26.txt | tail -12
