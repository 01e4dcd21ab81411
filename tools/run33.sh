#!/bin/bash
mkdir -p gpurun_out
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest33.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest33.log
tail -3 gpurun_out/pytest33.log
