#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/accumulate_probe.py 22,24 0 "128,12,4,48;128,12,2,48;128,12,1,48;128,12,3,48" > gpurun_out/ba_q21.txt 2>&1
grep -E "^---|mode=" gpurun_out/ba_q21.txt | sed -E 's/ \| .*(b_accumulate[a-z_]*=[0-9.]+).*/ \1/' | cut -c1-120
