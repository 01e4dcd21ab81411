#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/ba_variants17.txt
for v in "" latey; do
  echo "=== variant '${v:-default}'" >> gpurun_out/ba_variants17.txt
  SNARKV_LIB_VARIANT=$v timeout 300 python tools/accumulate_probe.py 22,23,24 0 >> gpurun_out/ba_variants17.txt 2>&1
done
grep -E "===|mode=2|equal" gpurun_out/ba_variants17.txt | cut -c1-250
SNARKV_LIB_VARIANT=latey timeout 600 python -m pytest tests/test_gpu_batched_affine.py -x -q -m gpu 2>&1 | tail -3
