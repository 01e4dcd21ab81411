#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_poseidon.py tests/test_plonk_verifier.py -x -q -m gpu > gpurun_out/pytest6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest6.log
tail -25 gpurun_out/pytest6.log
