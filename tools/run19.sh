#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest19.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest19.log
tail -3 gpurun_out/pytest19.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-aux --no-cpu-baseline > gpurun_out/bench19.json 2> gpurun_out/bench19.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench19.err
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/bench19.json") if l.startswith("{")][-1])
print("value %.1f ms %.3f e2e %.1f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), d["e2e"].get("host_numa"), d["stages_ms"])
PY
timeout 300 python tools/size_sweep.py 2>&1 | tail -22
