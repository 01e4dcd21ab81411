"""Generate tests/golden/*.json from the independent Python big-int model (oracle/bn254_model.py).

Run:  python oracle/gen_golden.py          (takes ~1 minute; output is committed)

The reference ships no known-answer vectors for this path ("parity unpinned"), so these fixtures are
produced by the Python model and asserted against BOTH the C++ oracle (tests, -m "not gpu") and the CUDA path
(tests, -m gpu).  All values are hex strings of the little-endian byte encodings fixed in include/snarkv_cuda.h.
"""
import json
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import bn254_model as m  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
le = m.fe_to_le


def hx(b):
    return bytes(b).hex()


def field_vectors():
    rnd = random.Random(101)
    out = {"p": hx(le(m.P)), "r": hx(le(m.R))}
    for name, mod in (("fq", m.P), ("fr", m.R)):
        R256 = 2**256 % mod
        out[name] = {
            "mont_R": hx(le(R256)), "mont_R2": hx(le(R256 * R256 % mod)),
            "inv64": "%016x" % ((-pow(mod, -1, 2**64)) % 2**64),
            "mul": [], "inv": [],
        }
        cases = [(0, 0), (1, 1), (mod - 1, mod - 1), (mod - 1, 2), (R256, R256)]
        cases += [(rnd.randrange(mod), rnd.randrange(mod)) for _ in range(32)]
        for a, b in cases:
            out[name]["mul"].append([hx(le(a)), hx(le(b)), hx(le(a * b % mod))])
        for a in [1, 2, mod - 1] + [rnd.randrange(1, mod) for _ in range(8)]:
            out[name]["inv"].append([hx(le(a)), hx(le(pow(a, -1, mod)))])
    return out


def g1_vectors():
    rnd = random.Random(202)
    ks = [0, 1, 2, 3, m.R - 1, m.R - 2, 2**253, 2**253 + 12345] + [rnd.randrange(m.R) for _ in range(12)]
    out = {"generator": hx(m.g1_to_bytes(m.G1_GEN)), "mul_G": [], "add": []}
    for k in ks:
        out["mul_G"].append([hx(le(k)), hx(m.g1_to_bytes(m.g1_mul(m.G1_GEN, k)))])
    pts = [m.g1_mul(m.G1_GEN, rnd.randrange(m.R)) for _ in range(6)]
    pairs = [(pts[0], pts[1]), (pts[2], pts[2]), (pts[3], m.g1_neg(pts[3])), (None, pts[4]), (pts[5], None), (None, None)]
    for a, b in pairs:
        out["add"].append([hx(m.g1_to_bytes(a)), hx(m.g1_to_bytes(b)), hx(m.g1_to_bytes(m.g1_add(a, b)))])
    return out


def msm_case(name, scalars, points):
    return {"name": name, "n": len(scalars),
            "scalars": hx(b"".join(le(s) for s in scalars)),
            "points": hx(b"".join(m.g1_to_bytes(p) for p in points)),
            "expected": hx(m.g1_to_bytes(m.msm_naive(scalars, points)))}


def msm_vectors():
    cases = []
    rnd = random.Random(303)
    def synth(seed, n):
        return ([m.synth_scalar(seed, i) for i in range(n)],
                [m.g1_mul(m.G1_GEN, m.synth_point_scalar(seed, i)) for i in range(n)])
    for n in (1, 2, 3, 21, 255, 256):
        s, p = synth(1000 + n, n)
        cases.append(msm_case(f"synth_n{n}", s, p))
    # edge cases the domain has: zero scalars, scalar one (Msm::base, util/msm.rs:54-61), identity points,
    # repeated bases, p and -p cancelling, scalars >= 2^253, all-equal scalars (single bucket)
    s, p = synth(7, 24)
    s[0] = 0; s[1] = 1; s[2] = m.R - 1; s[3] = 2**253 + 5; s[4] = 0
    p[5] = None; p[6] = None
    p[8] = p[7]; p[10] = m.g1_neg(p[9]); s[10] = s[9]
    cases.append(msm_case("edge_mixed_n24", s, p))
    s, p = synth(8, 40)
    cases.append(msm_case("all_ones_n40", [1] * 40, p))
    cases.append(msm_case("all_equal_scalar_n40", [s[0]] * 40, p))
    cases.append(msm_case("same_base_n17", s[:17], [p[0]] * 17))
    cases.append(msm_case("all_zero_scalars_n5", [0] * 5, p[:5]))
    cases.append(msm_case("all_identity_points_n5", s[:5], [None] * 5))
    q = p[3]
    cases.append(msm_case("cancels_to_identity_n2", [s[1], m.R - s[1]], [q, q]))
    # KzgAs::verify shape: powers of r (pcs/kzg/accumulation.rs:53-60)
    r = rnd.randrange(m.R)
    cases.append(msm_case("powers_of_r_n32", [pow(r, i, m.R) for i in range(32)], p[:32]))
    return cases


def pairing_vectors():
    out = {"g2_generator": hx(m.g2_to_bytes(m.G2_GEN))}
    out["e_G1_G2"] = hx(m.gt_to_bytes(m.pairing(m.G1_GEN, m.G2_GEN)))
    s = 0x0123456789ABCDEF0123456789ABCDEF0123456789ABCDEF01234567 % m.R
    s_g2 = m.g2_mul(m.G2_GEN, s)
    out["s"] = hx(le(s)); out["s_g2"] = hx(m.g2_to_bytes(s_g2))
    rnd = random.Random(404)
    checks = []
    def add(name, lhs, rhs):
        ok, gt = m.kzg_decide(lhs, rhs, m.G2_GEN, s_g2)
        checks.append({"name": name, "lhs": hx(m.g1_to_bytes(lhs)), "rhs": hx(m.g1_to_bytes(rhs)),
                       "accept": bool(ok), "gt": hx(m.gt_to_bytes(gt))})
    # the reference's own mock accumulator shape (system/halo2/test/kzg.rs:37-45): (lhs, rhs) = (s*G, G)
    add("mock_sG_G", m.g1_mul(m.G1_GEN, s), m.G1_GEN)
    for i in range(3):
        a = rnd.randrange(1, m.R)
        add(f"valid_{i}", m.g1_mul(m.G1_GEN, a * s % m.R), m.g1_mul(m.G1_GEN, a))
    a = rnd.randrange(1, m.R)
    add("tampered_rhs", m.g1_mul(m.G1_GEN, a * s % m.R), m.g1_mul(m.G1_GEN, a ^ 1))
    add("tampered_lhs", m.g1_mul(m.G1_GEN, (a * s + 1) % m.R), m.g1_mul(m.G1_GEN, a))
    add("both_identity", None, None)
    add("lhs_identity_only", None, m.g1_mul(m.G1_GEN, a))
    out["checks"] = checks
    return out


def accumulate_vectors():
    rnd = random.Random(505)
    s = 0x0123456789ABCDEF0123456789ABCDEF0123456789ABCDEF01234567 % m.R
    out = []
    for n in (1, 2, 3, 16):
        accs = []
        for _ in range(n):
            a = rnd.randrange(1, m.R)
            accs.append((m.g1_mul(m.G1_GEN, a * s % m.R), m.g1_mul(m.G1_GEN, a)))
        r = rnd.randrange(m.R)
        lhs, rhs = m.kzg_accumulate(accs, r)
        out.append({"n": n, "r": hx(le(r)),
                    "lhs": hx(b"".join(m.g1_to_bytes(a[0]) for a in accs)),
                    "rhs": hx(b"".join(m.g1_to_bytes(a[1]) for a in accs)),
                    "out_lhs": hx(m.g1_to_bytes(lhs)), "out_rhs": hx(m.g1_to_bytes(rhs))})
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    for name, fn in (("field", field_vectors), ("g1", g1_vectors), ("msm", msm_vectors),
                     ("pairing", pairing_vectors), ("accumulate", accumulate_vectors)):
        with open(os.path.join(OUT, name + ".json"), "w") as f:
            json.dump(fn(), f, indent=1)
        print("wrote", name)


if __name__ == "__main__":
    main()
