#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_plonk_verifier.py tests/test_pcs_mirror.py tests/test_plonk_eval.py -x -q -m gpu > gpurun_out/pytest4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest4.log
tail -15 gpurun_out/pytest4.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-sweep --no-cpu-baseline > gpurun_out/bench4.json 2> gpurun_out/bench4.err; echo "bench rc=$?"
tail -c 600 gpurun_out/bench4.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench4.json') if l.startswith('{')][-1])
for k,v in d.get('aux',{}).items():
    print(k, {x:v[x] for x in v if x in ('ms','proofs_per_s','checks_per_s','jobs_per_s','ok','error','program_instructions','lhs_terms_per_proof')})
PY
