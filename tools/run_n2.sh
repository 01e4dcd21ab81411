#!/bin/bash
# gpurun --gpus N payload: multi-GPU tests + value/e2e at N ranks (no aux, no CPU arm)
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_multirank_nccl.py -x -q > gpurun_out/pytest_n$N.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_n$N.log
tail -3 gpurun_out/pytest_n$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $N --steps 10 --warmup 3 --no-aux --no-cpu-baseline > gpurun_out/r02_bench_n${N}_noaux.json 2> gpurun_out/r02_bench_n${N}_noaux.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/r02_bench_n${N}_noaux.json") if l.startswith("{")][-1])
print("N=$N value %.1f ms %.3f e2e %.1f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), d["e2e"].get("host_numa"), d["stages_ms"])
PY
