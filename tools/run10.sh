#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_kzg_decide_fast -s 2 -c 1 -f -o gpurun_out/pf_n1 python tools/pairing_fast_ncu_target.py 1 > gpurun_out/pf_ncu.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/pf_ncu.log; ls -la gpurun_out/pf_n1.ncu-rep
