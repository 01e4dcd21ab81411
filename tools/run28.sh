#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest28.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest28.log
tail -3 gpurun_out/pytest28.log
timeout 300 python tools/accumulate_probe.py 21,22,24 0 > gpurun_out/ba_sqr28.txt 2>&1
grep -E "mode=|equal" gpurun_out/ba_sqr28.txt | sed -E 's/ \| .*(b_accumulate[a-z_]*=[0-9.]+).*(final=[0-9.]+).*/ \1 \2/' | cut -c1-140
