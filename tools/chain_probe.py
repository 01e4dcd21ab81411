"""Developer probe (not part of the bench contract): A/B/C of the bucket-accumulation kernels — accumulate mode 1 = XYZZ,
2 = batched-affine tree, 4 = chained batched affine (SNARKV_BC_R running sums per lane) — with stage timings.
usage: chain_probe.py [log2 sizes, comma separated] [window bits or 0] [R values, comma separated] [glv mode]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import snark_verifier_b200 as sv

sizes = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "20,22,24").split(",")]
cbits = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rs = [int(x) for x in (sys.argv[3] if len(sys.argv) > 3 else "16").split(",")]
glv = int(sys.argv[4]) if len(sys.argv) > 4 else 0
nmax = 1 << max(sizes)
stream = torch.cuda.Stream()
L = sv.CudaLoader(0)
L.set_stream(stream.cuda_stream)
with torch.cuda.stream(stream):
    ds = torch.empty(nmax * 32, dtype=torch.uint8, device="cuda")
    dp = torch.empty(nmax * 64, dtype=torch.uint8, device="cuda")
    out = torch.zeros(64 * 8, dtype=torch.uint8, device="cuda")
    L.synth_scalars_device(5, 0, nmax, ds.data_ptr())
    L.synth_points_device(5, 0, nmax, dp.data_ptr())
stream.synchronize()


def run(L, n, mode, slot, tag=""):
    L.set_accumulate_mode(mode)
    L.set_window_bits(cbits)
    L.set_glv_mode(glv)
    best = None
    for rep in range(4):
        L.profile(rep == 3)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            L.msm_device(ds.data_ptr(), dp.data_ptr(), n, d_out_affine=out.data_ptr() + 64 * slot)
            e1.record(stream)
        stream.synchronize()
        if 0 < rep < 3:
            t = e0.elapsed_time(e1)
            best = t if best is None else min(best, t)
    st = L.stage_times()
    acc = sum(b for a, b, _ in st if "accumulate" in a)
    print("n=2^%d c=%d mode=%d%s best %.3f ms %.1f Mterm/s accumulate %.3f ms | " % (n.bit_length() - 1, L.msm_plan(n)["window_bits"], mode, tag, best,
          n / best / 1e3, acc) + " ".join("%s=%.3f" % (a.replace("msm_", "").replace("bucket_", "b_"), b) for a, b, _ in st), flush=True)
    return best


for lg in sizes:
    n = 1 << lg
    run(L, n, 1, 0)
    run(L, n, 2, 1)
    for k, r in enumerate(rs):
        os.environ["SNARKV_BC_R"] = str(r)
        L2 = sv.CudaLoader(0)
        L2.set_stream(stream.cuda_stream)
        run(L2, n, 4, 2 + k, " R=%d" % r)
        L2.close()
    o = out.cpu().numpy()
    ref = bytes(o[:64])
    print("n=2^%d results equal: %s" % (lg, [bytes(o[64 * j:64 * j + 64]) == ref for j in range(1, 2 + len(rs))]), flush=True)
