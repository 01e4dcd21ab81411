"""Developer probe: end-to-end host-entry MSM (snarkv_g1_msm, pinned host buffers) under different developer knobs.
usage: e2e_probe.py log2n "ENV=VAL,ENV=VAL;ENV=VAL;..."   (each ';'-separated group is one run; the library reads SNARKV_* at init)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import snark_verifier_b200 as sv

lg = int(sys.argv[1])
groups = sys.argv[2].split(";") if len(sys.argv) > 2 else [""]
n = 1 << lg
L = sv.CudaLoader(0)
ds = torch.empty(n * 32, dtype=torch.uint8, device="cuda")
dp = torch.empty(n * 64, dtype=torch.uint8, device="cuda")
L.synth_scalars_device(5, 0, n, ds.data_ptr())
L.synth_points_device(5, 0, n, dp.data_ptr())
h_s = torch.empty(n * 32, dtype=torch.uint8).pin_memory()
h_p = torch.empty(n * 64, dtype=torch.uint8).pin_memory()
h_s.copy_(ds); h_p.copy_(dp)
torch.cuda.synchronize()
L.close()
ref = None
for g in groups:
    for kv in filter(None, g.split(",")):
        k, v = kv.split("=")
        os.environ[k] = v
    L = sv.CudaLoader(0)
    for _ in range(2):
        out = L.msm(h_s.numpy(), h_p.numpy(), n)
    ts = []
    for _ in range(4):
        t0 = time.perf_counter()
        out = L.msm(h_s.numpy(), h_p.numpy(), n)
        ts.append((time.perf_counter() - t0) * 1e3)
    ref = ref or out
    print("e2e n=2^%d [%s] best %.2f ms median %.2f ms  %.1f Mterm/s  same=%s" % (lg, g, min(ts), sorted(ts)[len(ts) // 2], n / min(ts) / 1e3, out == ref), flush=True)
    L.close()
