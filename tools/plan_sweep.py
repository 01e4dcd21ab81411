"""Developer probe: total MSM time over (window bits, GLV on/off, accumulate kernel) at the per-rank sizes of the multi-GPU runs.
usage: plan_sweep.py [log2 sizes] [window bits list] — prints one line per configuration, best of 3 (CUDA events)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import snark_verifier_b200 as sv

sizes = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "20,21,22").split(",")]
cs = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "13,14,15,16,17,18").split(",")]
nmax = 1 << max(sizes)
stream = torch.cuda.Stream()
L = sv.CudaLoader(0, fmt=sv.MONTGOMERY)
L.set_stream(stream.cuda_stream)
with torch.cuda.stream(stream):
    ds = torch.empty(nmax * 32, dtype=torch.uint8, device="cuda")
    dp = torch.empty(nmax * 64, dtype=torch.uint8, device="cuda")
    out = torch.zeros(64, dtype=torch.uint8, device="cuda")
    L.synth_scalars_device(5, 0, nmax, ds.data_ptr())
    L.synth_points_device(5, 0, nmax, dp.data_ptr())
stream.synchronize()
ref = {}
for lg in sizes:
    n = 1 << lg
    rows = []
    for glv in (1, 2):
        for c in cs:
            for mode in (1, 2):
                L.set_window_bits(c); L.set_glv_mode(glv); L.set_accumulate_mode(mode)
                best = None
                for rep in range(4):
                    L.profile(rep == 3)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    with torch.cuda.stream(stream):
                        e0.record(stream)
                        L.msm_device(ds.data_ptr(), dp.data_ptr(), n, d_out_affine=out.data_ptr())
                        e1.record(stream)
                    stream.synchronize()
                    if rep:
                        t = e0.elapsed_time(e1)
                        best = t if best is None else min(best, t)
                o = bytes(out.cpu().numpy())
                ref.setdefault(lg, o)
                st = {a.replace("msm_", ""): b for a, b, _ in L.stage_times()}
                acc = sum(v for k, v in st.items() if "accumulate" in k)
                tail = sum(v for k, v in st.items() if k in ("bucket_reduce", "window_sum", "final", "bucket_merge"))
                rows.append((best, glv, c, mode, acc, tail, o == ref[lg]))
    rows.sort()
    for best, glv, c, mode, acc, tail, ok in rows[:12]:
        print("n=2^%d glv=%s c=%2d mode=%d total %.3f ms  accumulate %.3f  tail %.3f  ok=%s" % (lg, "on " if glv == 1 else "off", c, mode, best, acc, tail, ok), flush=True)
    L.set_window_bits(0); L.set_glv_mode(0); L.set_accumulate_mode(0)
