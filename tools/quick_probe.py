"""Developer probe (not part of the bench contract): window-size sweep + stage timings of the MSM / decide pipelines."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import snark_verifier_b200 as sv

L = sv.CudaLoader(0)
stream = torch.cuda.Stream()
L.set_stream(stream.cuda_stream)
sizes = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "16,20,24").split(",")]
span = int(sys.argv[2]) if len(sys.argv) > 2 else 2
nmax = 1 << max(sizes)
with torch.cuda.stream(stream):
    ds = torch.empty(nmax * 32, dtype=torch.uint8, device="cuda")
    dp = torch.empty(nmax * 64, dtype=torch.uint8, device="cuda")
    out = torch.zeros(64, dtype=torch.uint8, device="cuda")
    t0 = time.time()
    L.synth_scalars_device(5, 0, nmax, ds.data_ptr()); L.synth_points_device(5, 0, nmax, dp.data_ptr())
stream.synchronize()
print("synth %d terms: %.3f s" % (nmax, time.time() - t0), flush=True)
for lg in sizes:
    n = 1 << lg
    L.set_window_bits(0)
    c0 = L.msm_plan(n)["window_bits"]
    for c in range(max(2, c0 - span), min(22, c0 + span) + 1):
        L.set_window_bits(c)
        best = None
        for rep in range(3):
            L.profile(rep == 2)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(stream):
                e0.record(stream); L.msm_device(ds.data_ptr(), dp.data_ptr(), n, d_out_affine=out.data_ptr()); e1.record(stream)
            stream.synchronize()
            if rep < 2:
                t = e0.elapsed_time(e1); best = t if best is None else min(best, t)
        st = L.stage_times()
        print("n=2^%d c=%d%s best %.3f ms %.1f Mterm/s | " % (lg, c, "*" if c == c0 else " ", best, n / best / 1e3) +
              " ".join("%s=%.3f" % (a.replace("msm_", "").replace("bucket_", "b_").replace("digits_", "d_"), b) for a, b, _ in st), flush=True)

# pairing probe: both kernels
import oracle
from oracle import bn254_model as m
g2 = oracle.g2_generator(); sk = 0x1234567; s_g2 = oracle.g2_mul(g2, m.fe_to_le(sk)); gen = m.g1_to_bytes(m.G1_GEN)
kz = sv.KzgAs(L, sv.KzgDecidingKey(gen, g2, s_g2))
L.profile(True)
for mode in (1, 2):
    L.set_pairing_mode(mode)
    for N in (1, 64, 1024, 4096, 1 << 14, 1 << 16):
        with torch.cuda.stream(stream):
            lhs = torch.frombuffer(bytearray(oracle.g1_mul(gen, m.fe_to_le(77 * sk)) * N), dtype=torch.uint8).cuda()
            rhs = torch.frombuffer(bytearray(oracle.g1_mul(gen, m.fe_to_le(77)) * N), dtype=torch.uint8).cuda()
            acc = torch.zeros(N, dtype=torch.uint8, device="cuda")
            for rep in range(2):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream); kz.decide_batch_device(lhs.data_ptr(), rhs.data_ptr(), N, acc.data_ptr()); e1.record(stream)
                stream.synchronize()
        print("decide mode=%d N=%d: %.3f ms  %.0f checks/s  all_accept=%s" % (mode, N, e0.elapsed_time(e1), N / e0.elapsed_time(e1) * 1e3, bool(acc.min().item() == 1)), flush=True)
