import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import snark_verifier_b200 as sv
L = sv.CudaLoader(0)
g2 = bytes.fromhex(
    "edf692d95cbdde46ddda5ef7d422436779445c5e66006a42761e1f12efde0018c212f3aeb785e49712e7a9353349aaf1255dfb31b7bf60723a480d9293938e19"
    "aa7dfa6601cce64c7bd3430c69e7d1e38f40cb8d8071ab4aeb6d8cdba55ec8125b9722d1dcdaac55f38eb37033314bbc95330c69ad999eec75f05f58d0890609")
gen = (1).to_bytes(32, "little") + (2).to_bytes(32, "little")
kz = sv.KzgAs(L, sv.KzgDecidingKey(gen, g2, g2))
n = 1 << 17
pts = torch.empty(n * 64, dtype=torch.uint8, device="cuda"); acc = torch.zeros(n, dtype=torch.uint8, device="cuda")
L.synth_points_device(7, 0, n, pts.data_ptr())
L.profile(True)
for mode, name in ((5, "fast"), (1, "thread"), (3, "block"), (4, "warp"), (0, "auto")):
    L.set_pairing_mode(mode)
    for cnt in (1, 148, 296, 444, 1024, 2368, 4096, 1 << 14, 1 << 16, 1 << 17):
        if mode in (3, 5) and cnt > 16384: continue
        if mode == 3 and cnt > 4096: continue
        for _ in range(2):
            kz.decide_batch_device(pts.data_ptr(), pts.data_ptr(), cnt, acc.data_ptr()); torch.cuda.synchronize()
        st = L.stage_times(); tot = sum(x[1] for x in st)
        print("%-6s N=%-6d %8.2f ms %9.0f checks/s ok=%s" % (name, cnt, tot, cnt / tot * 1e3, bool(acc[:cnt].min().item() == 1)), flush=True)
