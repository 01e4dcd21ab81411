#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/plan_sweep.py 16,17,18,19,20,21,22,23 12,13,14,15,16,17,18 > gpurun_out/plan_sweep18.txt 2>&1; cat gpurun_out/plan_sweep18.txt | awk '{k=$1; c[k]++; if (c[k]<=5) print}'
timeout 300 python tools/accumulate_probe.py 20,24 0 2>&1 | grep -E "mode=2|equal" | cut -c1-200
