#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_external_kat.py -x -q -m gpu -k "decide or pairing or latency or eip197 or montgomery_layout" > gpurun_out/pytest9.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest9.log
tail -15 gpurun_out/pytest9.log
timeout 300 python tools/pairing_probe.py > gpurun_out/pairing_probe9.txt 2>&1; cat gpurun_out/pairing_probe9.txt | head -40
