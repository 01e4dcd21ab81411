#!/bin/bash
# gpurun --gpus N payload: multi-GPU tests + bench at N ranks (torchrun) + the host chunk-pipeline A/B
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_n$N.txt
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_multirank_nccl.py -x -q > gpurun_out/pytest_n$N.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_n$N.log
tail -3 gpurun_out/pytest_n$N.log
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:2}"; }
run 29555 --steps 10 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo "bench rc=$?"
tail -c 300 gpurun_out/r02_bench_n$N.err
SNARKV_HOST_CHUNK_MIN=20 run 29556 --steps 10 --warmup 3 --no-aux --no-cpu-baseline > gpurun_out/r02_bench_n${N}_chunkmin20.json 2>/dev/null
SNARKV_HOST_CHUNK_MIN=21 SNARKV_HOST_CHUNKS_SMALL=2 run 29557 --steps 10 --warmup 3 --no-aux --no-cpu-baseline > gpurun_out/r02_bench_n${N}_chunkmin21x2.json 2>/dev/null
python - <<PY
import json
for f in ("r02_bench_n$N.json", "r02_bench_n${N}_chunkmin20.json", "r02_bench_n${N}_chunkmin21x2.json"):
    try:
        d = json.loads([l for l in open("gpurun_out/" + f) if l.startswith("{")][-1])
        print(f, "value %.1f ms %.2f e2e %.1f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
    except Exception as e:
        print(f, "failed", e)
PY
