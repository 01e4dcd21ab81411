#!/bin/bash
# gpurun --gpus N payload: value/e2e at N ranks only (no tests, no aux, no CPU arm)
N=${1:-8}
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus $N --steps 10 --warmup 3 --no-aux --no-cpu-baseline > gpurun_out/r02_bench_n${N}_noaux.json 2> gpurun_out/r02_bench_n${N}_noaux.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/r02_bench_n${N}_noaux.json") if l.startswith("{")][-1])
print("N=$N value %.1f ms %.3f e2e %.1f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), d["stages_ms"])
PY
