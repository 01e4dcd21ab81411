"""Developer probe: device-resident MSM time + per-stage CUDA-event times under different developer knobs.
usage: stage_probe.py 20,22,24 "ENV=VAL,ENV=VAL;ENV=VAL;..."   (one run per ';' group; the library reads SNARKV_* at snarkv_init)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import snark_verifier_b200 as sv

sizes = [int(x) for x in sys.argv[1].split(",")]
groups = sys.argv[2].split(";") if len(sys.argv) > 2 else [""]
nmax = 1 << max(sizes)
stream = torch.cuda.Stream()
L = sv.CudaLoader(0)
L.set_stream(stream.cuda_stream)
with torch.cuda.stream(stream):
    ds = torch.empty(nmax * 32, dtype=torch.uint8, device="cuda")
    dp = torch.empty(nmax * 64, dtype=torch.uint8, device="cuda")
    out = torch.zeros(64, dtype=torch.uint8, device="cuda")
    L.synth_scalars_device(5, 0, nmax, ds.data_ptr())
    L.synth_points_device(5, 0, nmax, dp.data_ptr())
stream.synchronize()
L.close()
ref = {}
for g in groups:
    for kv in filter(None, g.split(",")):
        k, v = kv.split("=")
        os.environ[k] = v
    L = sv.CudaLoader(0)
    L.set_stream(stream.cuda_stream)
    print("--- [%s]" % g, flush=True)
    for lg in sizes:
        n = 1 << lg
        best = None
        for rep in range(5):
            L.profile(rep == 4)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(stream):
                e0.record(stream)
                L.msm_device(ds.data_ptr(), dp.data_ptr(), n, d_out_affine=out.data_ptr())
                e1.record(stream)
            stream.synchronize()
            if 0 < rep < 4:
                t = e0.elapsed_time(e1)
                best = t if best is None else min(best, t)
        st = L.stage_times()
        res = bytes(out.cpu().numpy())
        ref.setdefault(lg, res)
        print("n=2^%d c=%d best %.3f ms %.1f Mterm/s same=%s | " % (lg, L.msm_plan(n)["window_bits"], best, n / best / 1e3, res == ref[lg]) +
              " ".join("%s=%.3f" % (a.replace("msm_", "").replace("bucket_", "b_").replace("digits_", "d_"), b) for a, b, _ in st), flush=True)
    L.close()
