#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/plonk_batch_stages.py > gpurun_out/plonk_batch_stages.txt 2>&1; cat gpurun_out/plonk_batch_stages.txt | tail -8
