#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_plonk_verifier.py tests/test_gpu_poseidon.py tests/test_pcs_mirror.py tests/test_pcs_boundary.py -x -q -m gpu > gpurun_out/pytest12.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest12.log
tail -5 gpurun_out/pytest12.log
timeout 300 python tools/plonk_batch_stages.py > gpurun_out/plonk_batch_stages2.txt 2>&1; cut -c1-120 gpurun_out/plonk_batch_stages2.txt | tail -4
SNARKV_OVERLAP_MSMS=0 timeout 300 python tools/plonk_batch_stages.py > gpurun_out/plonk_batch_stages2_serial.txt 2>&1; cut -c1-120 gpurun_out/plonk_batch_stages2_serial.txt | tail -4
