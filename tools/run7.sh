#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_plonk_verifier.py -x -q -m gpu > gpurun_out/pytest7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest7.log
tail -25 gpurun_out/pytest7.log
timeout 300 python tools/plonk_probe.py > gpurun_out/plonk_probe2.txt 2>&1; cat gpurun_out/plonk_probe2.txt | tail -5
