#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_external_kat.py tests/test_plonk_verifier.py -x -q -m gpu > gpurun_out/pytest32.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest32.log
tail -3 gpurun_out/pytest32.log
timeout 200 python tools/pairing_probe_thread.py 2>&1 | tail -4
