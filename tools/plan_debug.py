import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import snark_verifier_b200 as sv
from snark_verifier_b200 import plonk
import plonk_toy as T
srs = T.Srs(3); circuit = T.Circuit(4, 11, [5, 7])
L = sv.CudaLoader(0); kz = sv.KzgAs(L, sv.KzgDecidingKey(T.GEN, srs.g2, srs.s_g2))
for scheme in ("gwc19", "bdfg21"):
    protocol = T.make_protocol(circuit, srs, None)
    bv = plonk.PlonkBatchVerifier(L, kz, T.GEN, protocol, scheme)
    proof = T.prove(circuit, protocol, srs, scheme)
    inst = [circuit.public]
    plan_acc = bv.accumulate_new([inst], [proof], 1)
    bv.use_device_plan = False
    py_acc = bv.accumulate_new([inst], [proof], 1)
    print(scheme, "plan == python:", (plan_acc.lhs, plan_acc.rhs) == (py_acc.lhs, py_acc.rhs), "lhs eq", plan_acc.lhs == py_acc.lhs, "rhs eq", plan_acc.rhs == py_acc.rhs)
    rows, lookup, ch = bv.read_proofs([inst], [proof])
    prog = bv.compiled.msm.program
    lay = bv.compiled.layout
    def acc_from_rows(rows):
        out = np.frombuffer(L.fr_program_eval(prog, rows.tobytes(), 1), dtype=np.uint8).reshape(1, len(prog.outputs), 32)
        nl = len(bv._slots["lhs"]); res = []
        for side, sc in (("lhs", out[:, :nl]), ("rhs", out[:, nl:])):
            pts = bv._points(lookup, 1, side)
            off = np.arange(2, dtype=np.uint64) * pts.shape[1]
            res.append(L.msm_batch_rlc(np.ascontiguousarray(sc).reshape(-1), pts.reshape(-1), off, (1).to_bytes(32, "little")))
        return res
    base = acc_from_rows(rows)
    print("  recomputed python == python:", base[0] == py_acc.lhs)
    n = rows.shape[1]
    for i in range(n):
        for j in range(n):
            if i == j: continue
            r2 = rows.copy(); r2[0, i] = rows[0, j]
            a = acc_from_rows(r2)
            if a[0] == plan_acc.lhs:
                print("  MATCH when row[%d] := row[%d]" % (i, j), "layout", lay)
    # also: zero each row
    for i in range(n):
        r2 = rows.copy(); r2[0, i] = 0
        a = acc_from_rows(r2)
        if a[0] == plan_acc.lhs:
            print("  MATCH when row[%d] := 0" % i, lay)
    print("  layout", lay, "n_chal", bv.tl.n_challenges, "seg_end", bv.tl.seg_end)
