"""Developer probe (not part of the bench contract): A/B of the bucket-accumulation kernels (accumulate mode 1 = XYZZ,
2 = batched affine) with stage timings.
usage: accumulate_probe.py [log2 sizes, comma separated] [window bits or 0] ["K,PAIRS_MIN,Q,MIN_LOAD;..."]
The optional third argument sweeps the batched-affine tuning knobs (the library reads SNARKV_BA_* at snarkv_init)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import snark_verifier_b200 as sv

sizes = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "20,22,24").split(",")]
cbits = int(sys.argv[2]) if len(sys.argv) > 2 else 0
sweep = [tuple(x.split(",")) for x in sys.argv[3].split(";")] if len(sys.argv) > 3 else [None]
nmax = 1 << max(sizes)
stream = torch.cuda.Stream()
L = sv.CudaLoader(0)
L.set_stream(stream.cuda_stream)
with torch.cuda.stream(stream):
    ds = torch.empty(nmax * 32, dtype=torch.uint8, device="cuda")
    dp = torch.empty(nmax * 64, dtype=torch.uint8, device="cuda")
    out = torch.zeros(128, dtype=torch.uint8, device="cuda")
    L.synth_scalars_device(5, 0, nmax, ds.data_ptr())
    L.synth_points_device(5, 0, nmax, dp.data_ptr())
stream.synchronize()


def run(L, n, mode):
    L.set_accumulate_mode(mode)
    best = None
    for rep in range(4):
        L.profile(rep == 3)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            L.msm_device(ds.data_ptr(), dp.data_ptr(), n, d_out_affine=out.data_ptr() + 64 * (mode - 1))
            e1.record(stream)
        stream.synchronize()
        if 0 < rep < 3:
            t = e0.elapsed_time(e1)
            best = t if best is None else min(best, t)
    st = L.stage_times()
    print("n=2^%d c=%d mode=%d best %.3f ms %.1f Mterm/s | " % (n.bit_length() - 1, L.msm_plan(n)["window_bits"], mode, best, n / best / 1e3) +
          " ".join("%s=%.3f" % (a.replace("msm_", "").replace("bucket_", "b_").replace("digits_", "d_"), b) for a, b, _ in st), flush=True)
    return best


base = {}
for k, cfg in enumerate(sweep):
    if cfg is not None:
        os.environ["SNARKV_BA_K"], os.environ["SNARKV_BA_PAIRS_MIN"], os.environ["SNARKV_BA_Q"], os.environ["SNARKV_BA_MIN_LOAD"] = cfg
        L.close()
        L = sv.CudaLoader(0)
        L.set_stream(stream.cuda_stream)
        print("--- K,PAIRS_MIN,Q,MIN_LOAD = %s" % (cfg,), flush=True)
    L.set_window_bits(cbits)
    for lg in sizes:
        n = 1 << lg
        if k == 0:
            base[lg] = run(L, n, 1)
        t2 = run(L, n, 2)
        o = out.cpu().numpy()
        print("n=2^%d results equal: %s   speed-up over XYZZ %.3fx" % (lg, bytes(o[:64]) == bytes(o[64:]), base[lg] / t2), flush=True)
