#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/accumulate_probe.py 23,24 0 "128,24,4,48;128,12,4,48;128,6,4,48;128,3,4,48;64,6,4,48" > gpurun_out/ba_pairs_min14.txt 2>&1; cat gpurun_out/ba_pairs_min14.txt
