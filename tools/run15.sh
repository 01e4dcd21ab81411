#!/bin/bash
# A/B of batched-affine kernel build variants (make variant V=...): forward prefetch depth 2, inlined backward multiplications
mkdir -p gpurun_out
: > gpurun_out/ba_variants15.txt
for v in "" fwd2 inl inlfwd2; do
  echo "=== variant '${v:-default}'" >> gpurun_out/ba_variants15.txt
  SNARKV_LIB_VARIANT=$v timeout 300 python tools/accumulate_probe.py 22,23,24 0 >> gpurun_out/ba_variants15.txt 2>&1
done
grep -E "===|mode=2|equal" gpurun_out/ba_variants15.txt | cut -c1-110
