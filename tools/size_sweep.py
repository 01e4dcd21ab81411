"""BASELINE.json config 5, N = 1 leg: MSM size sweep 2^lo .. 2^hi with the library's own plan (window bits, GLV, accumulate
mode all automatic), operands resident in HBM, CUDA events on the launch stream; plus a 2^20-check KZG decide batch.
usage: size_sweep.py [lo] [hi] [pairing log2 N]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import snark_verifier_b200 as sv

lo = int(sys.argv[1]) if len(sys.argv) > 1 else 10
hi = int(sys.argv[2]) if len(sys.argv) > 2 else 26
plg = int(sys.argv[3]) if len(sys.argv) > 3 else 20
stream = torch.cuda.Stream()
L = sv.CudaLoader(0)
L.set_stream(stream.cuda_stream)
nmax = 1 << hi
with torch.cuda.stream(stream):
    ds = torch.empty(nmax * 32, dtype=torch.uint8, device="cuda")
    dp = torch.empty(nmax * 64, dtype=torch.uint8, device="cuda")
    out = torch.zeros(64, dtype=torch.uint8, device="cuda")
    L.synth_scalars_device(7, 0, nmax, ds.data_ptr())
    L.synth_points_device(7, 0, nmax, dp.data_ptr())
stream.synchronize()
print("# MSM size sweep, 1 x B200, operands in HBM, best of 3 after 2 warm-ups")
print("%6s %4s %8s %10s %12s  %s" % ("log2n", "c", "windows", "ms", "Mterm/s", "result (first 8 bytes of x)"))
for lg in range(lo, hi + 1):
    n = 1 << lg
    ts = []
    for rep in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            L.msm_device(ds.data_ptr(), dp.data_ptr(), n, d_out_affine=out.data_ptr())
            e1.record(stream)
        stream.synchronize()
        if rep >= 2:
            ts.append(e0.elapsed_time(e1))
    pl = L.msm_plan(n)
    print("%6d %4d %8d %10.3f %12.2f  %s" % (lg, pl["window_bits"], pl["windows"], min(ts), n / min(ts) / 1e3, bytes(out.cpu().numpy()[:8]).hex()), flush=True)

# ---- 2^plg independent KZG decisions (config 5's pairing batch): key with s = 1, so (P, P) accepts and (P, Q != P) rejects ----
g2 = bytes.fromhex(
    "edf692d95cbdde46ddda5ef7d422436779445c5e66006a42761e1f12efde0018c212f3aeb785e49712e7a9353349aaf1255dfb31b7bf60723a480d9293938e19"
    "aa7dfa6601cce64c7bd3430c69e7d1e38f40cb8d8071ab4aeb6d8cdba55ec8125b9722d1dcdaac55f38eb37033314bbc95330c69ad999eec75f05f58d0890609")
gen = (1).to_bytes(32, "little") + (2).to_bytes(32, "little")
kz = sv.KzgAs(L, sv.KzgDecidingKey(gen, g2, g2))
N = 1 << plg
with torch.cuda.stream(stream):
    acc = torch.zeros(N, dtype=torch.uint8, device="cuda")
    lhs = dp[: N * 64]
    rhs = lhs.clone()
    # every 7th check is made invalid: rhs_i <- the next point
    view_r = rhs.view(N, 64)
    view_l = lhs.view(N, 64)
    bad = torch.arange(0, N - 1, 7, device="cuda")
    view_r[bad] = view_l[bad + 1]
stream.synchronize()
ts = []
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        kz.decide_batch_device(lhs.data_ptr(), rhs.data_ptr(), N, acc.data_ptr())
        e1.record(stream)
    stream.synchronize()
    if rep:
        ts.append(e0.elapsed_time(e1))
a = acc.cpu()
expect = torch.ones(N, dtype=torch.uint8)
expect[bad.cpu()] = 0
print("# KZG decide batch: N=2^%d independent 2-pair checks (every 7th invalid): %.2f ms, %.0f checks/s, accept vector correct: %s"
      % (plg, min(ts), N / min(ts) * 1e3, bool((a == expect).all())))
L.close()
