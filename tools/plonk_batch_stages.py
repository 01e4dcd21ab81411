"""Developer probe: CUDA-event stage times inside ONE PlonkBatchVerifier.verify_batch call on the device-resident plan
(4096 fixture proofs), per scheme / transcript, next to the call's wall clock; then decide_all_fused and KzgAs::verify + decide."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import snark_verifier_b200 as sv
from snark_verifier_b200 import plonk
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
fx = json.load(open(os.path.join(ROOT, "tests", "golden", "plonk_proofs.json")))
H = bytes.fromhex
L = sv.CudaLoader(0)
kz = sv.KzgAs(L, sv.KzgDecidingKey(H(fx["svk_g"]), H(fx["g2"]), H(fx["s_g2"])))
protocol = plonk.simple_plonk_protocol(fx["k"], [H(p) for p in fx["preprocessed"]], fx["num_public"], None, fx["initial_state"])
m = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
for scheme, tr in (("gwc19", "evm"), ("bdfg21", "evm"), ("bdfg21", "poseidon")):
    bv = plonk.PlonkBatchVerifier(L, kz, H(fx["svk_g"]), protocol, scheme, transcript=tr)
    good = [e for e in fx[scheme if tr == "evm" else scheme + "_" + tr] if e["valid"]]
    order = "big" if tr == "evm" else "little"
    insts = [[[int(v) for v in col] for col in good[j % 8]["instances"]] for j in range(m)]
    proofs = np.frombuffer(b"".join(H(good[j % 8]["proof"]) for j in range(m)), dtype=np.uint8).reshape(m, -1)
    insts = np.frombuffer(b"".join(v.to_bytes(32, order) for inst in insts for col in inst for v in col), dtype=np.uint8).reshape(m, -1, 32)
    for _ in range(3):
        assert bv.verify_batch(insts, proofs, 12345) is True
    ts = []
    for _ in range(5):
        t0 = time.perf_counter(); bv.verify_batch(insts, proofs, 12345); ts.append((time.perf_counter() - t0) * 1e3)
    L.profile(True)
    bv.verify_batch(insts, proofs, 12345)
    st = L.stage_times()
    L.profile(False)
    print("%s %s m=%d wall best %.2f ms median %.2f ms; stages (ms x launches): %s ; sum %.2f" % (
        scheme, tr, m, min(ts), sorted(ts)[2], " ".join("%s=%.3f" % (n, t) for n, t, k in st), sum(t for _, t, _ in st)), flush=True)
    bv.close()
