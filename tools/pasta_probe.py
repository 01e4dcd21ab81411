"""Developer probe: Pallas MSM timing (operands resident, library's own plan) and IPA decide latency at 2^k."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import snark_verifier_b200 as sv
from snark_verifier_b200 import pasta
sizes = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "16,20,22,24").split(",")]
stream = torch.cuda.Stream()
L = sv.CudaLoader(0, fmt=sv.MONTGOMERY)
L.set_stream(stream.cuda_stream)
PL = pasta.PallasLoader(L)
nmax = 1 << max(sizes)
with torch.cuda.stream(stream):
    ds = torch.empty(nmax * 32, dtype=torch.uint8, device="cuda")
    dp = torch.empty(nmax * 64, dtype=torch.uint8, device="cuda")
    out = torch.zeros(64, dtype=torch.uint8, device="cuda")
    PL.synth_scalars_device(5, 0, nmax, ds.data_ptr())
    PL.synth_points_device(5, 0, nmax, dp.data_ptr())
stream.synchronize()
for lg in sizes:
    n = 1 << lg
    best = None
    for rep in range(4):
        L.profile(rep == 3)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            PL.msm_device(ds.data_ptr(), dp.data_ptr(), n, d_out_affine=out.data_ptr())
            e1.record(stream)
        stream.synchronize()
        if rep:
            t = e0.elapsed_time(e1)
            best = t if best is None else min(best, t)
    print("pallas msm n=2^%d best %.3f ms %.1f Mterm/s | " % (lg, best, n / best / 1e3) +
          " ".join("%s=%.3f" % (a.replace("msm_", "").replace("bucket_", "b_"), b) for a, b, _ in L.stage_times()), flush=True)
# IPA decide: key of 2^k points, one accumulator
Lc = sv.CudaLoader(0)
PLc = pasta.PallasLoader(Lc)
for k in (12, 16, 20):
    n = 1 << k
    dpc = torch.empty(n * 64, dtype=torch.uint8, device="cuda")
    PLc.synth_points_device(7, 0, n, dpc.data_ptr())
    torch.cuda.synchronize()
    g = dpc.cpu().numpy().tobytes()
    t0 = time.perf_counter()
    ipa = pasta.IpaAs(Lc, pasta.IpaDecidingKey(g))
    t1 = time.perf_counter()
    xi = [(3 + i).to_bytes(32, "little") for i in range(k)]
    h = PLc.h_coeffs(b"".join(xi), k)
    u = PLc.msm(h, g, n)
    acc = pasta.IpaAccumulator(xi, u)
    ipa.decide(acc)
    t2 = time.perf_counter()
    for _ in range(5):
        ipa.decide(acc)
    t3 = time.perf_counter()
    print("ipa k=%d: key upload+validate %.1f ms, decide %.3f ms (wall, incl. H2D of xi/u)" % (k, (t1 - t0) * 1e3, (t3 - t2) / 5 * 1e3), flush=True)
