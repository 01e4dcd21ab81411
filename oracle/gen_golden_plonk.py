"""Generate tests/golden/plonk_eval.json and tests/golden/limbs.json from the big-integer oracle (oracle/plonk_eval_model.py,
oracle/bn254_model.py).  TEST INFRASTRUCTURE.

Run:  python oracle/gen_golden_plonk.py     (seconds; output is committed)

plonk_eval.json pins SURVEY §8 row f3: a StandardPlonk-shaped quotient numerator written out here as plain nested tuples
(independently of snark_verifier_b200/plonk_eval.py, whose `standard_plonk_like_protocol(...).numerator.to_tuple()` must equal it),
per-proof input rows and the values the reference's evaluation yields for them: [quotient_eval, z^n, z^n - 1, 1/(z^n - 1),
instance evaluations].  limbs.json pins row a13: accumulators as 4 x 68-bit limbs and what `LimbsEncoding::from_repr` returns —
including rows it must reject.  "Parity unpinned" (no known-answer vectors in the reference); values are canonical mathematics.
"""
import json
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import bn254_model as m  # noqa: E402
import plonk_eval_model as om  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
R, P = om.R, m.P
DELTA = pow(7, 1 << 28, R)


def poly(i, r=0):
    return ("poly", i, r)


def add(a, b):
    return ("sum", a, b)


def mul(a, b):
    return ("product", a, b)


def sub(a, b):
    return ("sum", a, ("neg", b))


def standard_plonk_numerator(blinding_factors=5):
    """gate + permutation argument over (a, b, c) with zero-knowledge rows, as system/halo2.rs:520-660 lays the constraints out;
    polynomial indices: 0-4 fixed, 5-7 sigmas, 8 instance, 9-11 advice, 12 permutation z; challenges 1 beta, 2 gamma, 3 alpha."""
    q_a, q_b, q_c, q_ab, constant = (poly(i) for i in range(5))
    sigmas = [poly(5 + i) for i in range(3)]
    advice = [poly(9 + i) for i in range(3)]
    a, b, c = advice
    z, z_omega = poly(12), poly(12, 1)
    beta, gamma, alpha = ("challenge", 1), ("challenge", 2), ("challenge", 3)
    one = ("constant", 1)
    last = -(blinding_factors + 1)
    lag = lambda i: ("common", "lagrange", i)
    l_blind = None
    for i in range(last + 1, 0):
        l_blind = lag(i) if l_blind is None else add(l_blind, lag(i))
    l_active = sub(one, add(lag(last), l_blind))
    identity = ("common", "identity", 0)
    gate = add(add(add(add(add(mul(q_a, a), mul(q_b, b)), mul(q_c, c)), mul(mul(q_ab, a), b)), constant), poly(8))
    left = z_omega
    for p, s in zip(advice, sigmas):
        left = mul(left, add(add(p, mul(beta, s)), gamma))
    right, delta = z, 1
    for p in advice:
        right = mul(right, add(add(p, mul(mul(beta, ("constant", delta)), identity)), gamma))
        delta = delta * DELTA % R
    constraints = (gate, mul(lag(0), sub(one, z)), mul(lag(last), sub(mul(z, z), z)), mul(l_active, sub(left, right)))
    return ("powers", constraints, alpha)


def plonk_eval_fixture():
    rnd = random.Random(606)
    out = {"num_preprocessed": 8, "num_challenge": 4, "blinding_factors": 5,
           "evaluation_queries": [[i, 0] for i in range(8)] + [[9 + i, 0] for i in range(3)] + [[12, 0], [12, 1]],
           "input_layout": "z | 4 challenges | 13 evaluations (evaluation_queries order) | instances", "cases": []}
    num = standard_plonk_numerator()
    out["numerator"] = num
    for k, ninst in ((4, 1), (8, 3), (12, 2)):
        rows = []
        tot = 1 + 4 + 13 + ninst
        for j in range(5):
            row = [rnd.randrange(R) for _ in range(tot)]
            if j == 3:
                row[0] = om.root_of_unity(k)          # z on the domain: zero denominators stay zero
            if j == 4:
                row[0] = 0
            exp = om.quotient_evaluation(k, 8, [ninst], [tuple(q) for q in out["evaluation_queries"]], num, row[0], row[1:5], row[5:18], [row[18:]])
            rows.append({"inputs": ["%064x" % v for v in row], "outputs": ["%064x" % v for v in exp]})
        out["cases"].append({"k": k, "num_instance": ninst, "rows": rows})
    return out


def limbs_fixture():
    rnd = random.Random(707)
    pt = lambda: m.g1_mul(m.G1_GEN, rnd.randrange(1, m.R))
    to_limbs = lambda v: [(v >> (68 * i)) & ((1 << 68) - 1) for i in range(4)]
    rows = []
    for j in range(12):
        lhs, rhs = pt(), pt()
        if j == 2:
            lhs = None                                   # identity: (0, 0)
        coords = [c for p in (lhs, rhs) for c in ((0, 0) if p is None else p)]
        limbs = [l for c in coords for l in to_limbs(c)]
        note = "valid"
        if j == 5:
            limbs[0] ^= 1; note = "lhs off the curve"
        if j == 7:
            limbs[11] = (P >> 204) + 3; note = "rhs.x >= p"
        if j == 9:
            limbs[15] += 1 << 70; note = "rhs.y does not fit 32 bytes"
        if j == 10:
            limbs[4] += 1 << 68                          # limb wider than 68 bits but the sum is still a valid coordinate? no: changes y
            note = "lhs.y changed by an over-wide limb: off the curve"
        vals = [sum(l << (68 * i) for i, l in enumerate(limbs[4 * c:4 * c + 4])) for c in range(4)]
        ok = all(v < P for v in vals)
        if ok:
            for x, y in ((vals[0], vals[1]), (vals[2], vals[3])):
                ok = ok and ((x, y) == (0, 0) or (y * y - x * x * x - 3) % P == 0)
        rows.append({"note": note, "limbs": ["%064x" % l for l in limbs], "valid": int(ok),
                     "lhs": (m.fe_to_le(vals[0]) + m.fe_to_le(vals[1])).hex() if ok else "00" * 64,
                     "rhs": (m.fe_to_le(vals[2]) + m.fe_to_le(vals[3])).hex() if ok else "00" * 64})
    return {"limbs": 4, "bits": 68, "rows": rows}


if __name__ == "__main__":
    for name, fx in (("plonk_eval", plonk_eval_fixture()), ("limbs", limbs_fixture())):
        with open(os.path.join(OUT, name + ".json"), "w") as f:
            json.dump(fx, f, indent=0, separators=(",", ":"))
        print("wrote", name)
