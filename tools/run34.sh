#!/bin/bash
mkdir -p gpurun_out
timeout 100 python tools/size_sweep.py > gpurun_out/r02_size_sweep_n1.txt 2>&1; tail -21 gpurun_out/r02_size_sweep_n1.txt | cut -c1-160
