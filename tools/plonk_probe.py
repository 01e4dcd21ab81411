"""Developer probe: where the time of PlonkBatchVerifier.verify_batch goes (4096 fixture proofs), per transcript."""
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
m = 4096
for scheme, tr in (("gwc19", "evm"), ("bdfg21", "evm"), ("bdfg21", "poseidon")):
    bv = plonk.PlonkBatchVerifier(L, kz, H(fx["svk_g"]), protocol, scheme, transcript=tr)
    good = [e for e in fx[scheme if tr == "evm" else scheme + "_" + tr] if e["valid"]]
    insts = [[[int(v) for v in col] for col in good[j % 8]["instances"]] for j in range(m)]
    proofs = [H(good[j % 8]["proof"]) for j in range(m)]
    assert bv.verify_batch(insts, proofs, 12345) is True
    T = {}
    def timed(name, fn):
        t0 = time.perf_counter(); r = fn(); T[name] = T.get(name, 0) + (time.perf_counter() - t0) * 1e3; return r
    for rep in range(3):
        T.clear()
        st, words, lookup = timed("parse(+decompress)", lambda: bv._parse(insts, proofs))
        tl = bv.tl
        if tr == "evm":
            ch = timed("transcript", lambda: L.evm_transcript_challenges(st.tobytes(), tl.total * 32, [32 * e for e in tl.seg_end], m))
        else:
            ch = timed("transcript", lambda: L.poseidon_transcript_challenges(st.tobytes(), tl.total, list(tl.seg_end), m))
        rows, lookup, _ = timed("read_proofs(total)", lambda: bv.read_proofs(insts, proofs))
        prog = bv.compiled.msm.program
        out = timed("program", lambda: L.fr_program_eval(prog, rows.tobytes(), m))
        out = np.frombuffer(out, dtype=np.uint8).reshape(m, len(prog.outputs), 32)
        nl = len(bv._slots["lhs"])
        for side, sc in (("lhs", out[:, :nl]), ("rhs", out[:, nl:])):
            pts = timed("points_" + side, lambda: bv._points(lookup, m, side))
            off = np.arange(m + 1, dtype=np.uint64) * pts.shape[1]
            timed("msm_" + side, lambda: L.msm_batch_rlc(np.ascontiguousarray(sc).reshape(-1), pts.reshape(-1), off, (12345).to_bytes(32, "little"), flags=sv.CHECK_INPUTS))
        t0 = time.perf_counter(); ok = bv.verify_batch(insts, proofs, 12345); T["verify_batch(total)"] = (time.perf_counter() - t0) * 1e3
    print(scheme, tr, " ".join("%s=%.2f" % kv for kv in T.items()), flush=True)
