#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_poseidon.py tests/test_plonk_verifier.py -x -q -m gpu > gpurun_out/pytest8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest8.log
tail -12 gpurun_out/pytest8.log
timeout 300 python tools/plonk_probe.py > gpurun_out/plonk_probe3.txt 2>&1; cat gpurun_out/plonk_probe3.txt | tail -5
timeout 600 python bench.py --steps 3 --warmup 3 --no-sweep --no-cpu-baseline > gpurun_out/bench8.json 2> gpurun_out/bench8.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench8.json') if l.startswith('{')][-1])
for k,v in d.get('aux',{}).items():
    print(k, {x:v[x] for x in v if x in ('ms','proofs_per_s','checks_per_s','jobs_per_s','ok','error')})
PY
