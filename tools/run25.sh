#!/bin/bash
mkdir -p gpurun_out
S=$(date +%s)
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -x -q -m gpu -k "not full_size and not huge and not multirank and not 2p2" > gpurun_out/sanitizer_memcheck25.log 2>&1; echo "memcheck-all rc=$? in $(( $(date +%s) - S )) s"
tail -4 gpurun_out/sanitizer_memcheck25.log
S=$(date +%s)
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_plonk_verifier.py tests/test_gpu_external_kat.py tests/test_gpu_pasta.py -x -q -m gpu -k "not full_size" > gpurun_out/sanitizer_racecheck25.log 2>&1; echo "racecheck-b rc=$? in $(( $(date +%s) - S )) s"
tail -4 gpurun_out/sanitizer_racecheck25.log
