#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/ba_stagger23.txt
for sgr in 0 1; do
  echo "=== SNARKV_BA_STAGGER=$sgr" >> gpurun_out/ba_stagger23.txt
  SNARKV_BA_STAGGER=$sgr timeout 300 python tools/accumulate_probe.py 21,22,23,24 0 >> gpurun_out/ba_stagger23.txt 2>&1
done
grep -E "^===|mode=2|equal" gpurun_out/ba_stagger23.txt | sed -E 's/ \| .*(b_accumulate[a-z_]*=[0-9.]+).*/ \1/' | cut -c1-120
