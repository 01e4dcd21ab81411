#!/bin/bash
# FINAL N=1 pass: smoke + full GPU suite + bench + ncu launch list + ncu --set full capture (traffic JSON on the final kernel sources)
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest27.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest27.log
tail -3 gpurun_out/pytest27.log
timeout 900 python tools/ncu_traffic.py 24 gpurun_out > gpurun_out/r02_ncu_traffic.log 2>&1; echo "ncu full rc=$?"
cp gpurun_out/r02_traffic.json profiles/r02_traffic.json
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"
tail -c 300 gpurun_out/r02_bench_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-aux --no-cpu-baseline > gpurun_out/r02_launch_bench.log 2>&1; echo "ncu list rc=$?"
rm -f gpurun_out/*.ncu-rep.tmp
