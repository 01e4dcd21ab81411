#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/ba_phase22.txt
for v in "" skipfwd skipbwd skipinv skipfwdinv; do
  echo "=== variant '${v:-default}'" >> gpurun_out/ba_phase22.txt
  SNARKV_LIB_VARIANT=$v timeout 300 python tools/accumulate_probe.py 22,24 0 >> gpurun_out/ba_phase22.txt 2>&1
done
grep -E "^===|mode=2" gpurun_out/ba_phase22.txt | sed -E 's/ \| .*(b_accumulate[a-z_]*=[0-9.]+).*/ \1/' | cut -c1-120
