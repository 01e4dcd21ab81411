#!/bin/bash
# full GPU suite (timed) + per-rank plan sweep after the 8-window GLV plan
mkdir -p gpurun_out
S=$(date +%s)
timeout 1500 python -m pytest tests -x -q -m gpu --durations=12 > gpurun_out/pytest13.log 2>&1; echo "pytest rc=$? in $(( $(date +%s) - S )) s" >> gpurun_out/pytest13.log
tail -25 gpurun_out/pytest13.log
timeout 400 python tools/plan_sweep.py 20,21,22 15,16,17 > gpurun_out/plan_sweep13.txt 2>&1; cat gpurun_out/plan_sweep13.txt
