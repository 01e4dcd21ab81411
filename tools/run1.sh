#!/bin/bash
# gpurun payload: tests + quick bench + probes; everything lands in gpurun_out/
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -5 gpurun_out/pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
tail -c 600 gpurun_out/bench_n1.err
timeout 300 python tools/accumulate_probe.py 20,21,22,23 0 > gpurun_out/acc_probe.txt 2>&1
tail -12 gpurun_out/acc_probe.txt
