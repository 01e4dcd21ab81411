"""ncu target (developer tool): the latency decision kernel (pairing_fast.cu, mode 5) on N checks (argv[1], default 1)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import snark_verifier_b200 as sv
L = sv.CudaLoader(0)
g2 = bytes.fromhex(
    "edf692d95cbdde46ddda5ef7d422436779445c5e66006a42761e1f12efde0018c212f3aeb785e49712e7a9353349aaf1255dfb31b7bf60723a480d9293938e19"
    "aa7dfa6601cce64c7bd3430c69e7d1e38f40cb8d8071ab4aeb6d8cdba55ec8125b9722d1dcdaac55f38eb37033314bbc95330c69ad999eec75f05f58d0890609")
gen = (1).to_bytes(32, "little") + (2).to_bytes(32, "little")
kz = sv.KzgAs(L, sv.KzgDecidingKey(gen, g2, g2))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
pts = torch.empty(n * 64, dtype=torch.uint8, device="cuda")
acc = torch.zeros(n, dtype=torch.uint8, device="cuda")
L.synth_points_device(7, 0, n, pts.data_ptr())
L.set_pairing_mode(5)
for _ in range(3):
    kz.decide_batch_device(pts.data_ptr(), pts.data_ptr(), n, acc.data_ptr())
    torch.cuda.synchronize()
print("N", n, "all accept:", bool(acc[:n].min().item() == 1))
