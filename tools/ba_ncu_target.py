"""ncu target: a few MSMs of one size in a chosen accumulate mode (developer tool).  usage: ba_ncu_target.py log2n mode [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import snark_verifier_b200 as sv

lg, mode = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
n = 1 << lg
L = sv.CudaLoader(0)
ds = torch.empty(n * 32, dtype=torch.uint8, device="cuda")
dp = torch.empty(n * 64, dtype=torch.uint8, device="cuda")
out = torch.zeros(64, dtype=torch.uint8, device="cuda")
L.synth_scalars_device(5, 0, n, ds.data_ptr())
L.synth_points_device(5, 0, n, dp.data_ptr())
L.set_accumulate_mode(mode)
for _ in range(reps):
    L.msm_device(ds.data_ptr(), dp.data_ptr(), n, d_out_affine=out.data_ptr())
torch.cuda.synchronize()
print(bytes(out.cpu().numpy()).hex())
