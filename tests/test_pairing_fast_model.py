"""CPU: the thread-by-thread emulation of csrc/pairing_fast.cu's data flow (oracle/pairing_fast_model.py) against the
independent pairing model — pins the term map of the 192-thread Fq12 product, the Frobenius sign rule, the merged-line
slots, the signed-digit exponentiation, the norm inversion and the final-exponentiation chain."""
import random

from oracle import bn254_model as m
from oracle import pairing_fast_model as pf


def rnd12(rng):
    return [rng.randrange(m.P) for _ in range(12)]


def test_product_frobenius_conj_inverse():
    rng = random.Random(5)
    for _ in range(3):
        a, b = rnd12(rng), rnd12(rng)
        want = pf.w_from_model(m.f12_mul(pf.w_to_model(a), pf.w_to_model(b)))
        assert pf.op_mul(a, b) == want
        assert pf.op_mul(a, b, pf.xi_words(b)) == want
    a = rnd12(rng)
    for k in (1, 2, 3):
        assert pf.op_frobenius(a, k) == pf.w_from_model(m.f12_pow(pf.w_to_model(a), m.P ** k))
    assert pf.op_conj(a) == pf.w_from_model(m.f12_pow(pf.w_to_model(a), m.P ** 6))
    assert pf.op_mul(a, pf.op_inverse(a)) == [1] + [0] * 11


def test_u_digits_match_the_generated_table():
    import os, re
    inc = open(os.path.join(os.path.dirname(__file__), "..", "snark_verifier_b200", "csrc", "pairing_consts.inc")).read()
    digits = [int(x) for x in re.search(r"#define SNARKV_U_NAF_INIT \{([^}]*)\}", inc).group(1).split(",")]
    assert digits == pf.u_naf_msb_first()
    assert int(re.search(r"#define SNARKV_U_NAF_LEN (\d+)", inc).group(1)) == len(digits)
    v = 0
    for d in digits:
        v = 2 * v + d
    assert v == m.U


def test_schedule_decides_like_the_model():
    sk = 0x1234567
    g2 = m.G2_GEN
    s_g2 = m.g2_mul(g2, sk)
    G = m.G1_GEN
    cases = [(m.g1_mul(G, 77 * sk), m.g1_mul(G, 77)),          # accept
             (m.g1_mul(G, 77 * sk + 1), m.g1_mul(G, 77)),      # reject
             (None, m.g1_mul(G, 5)), (m.g1_mul(G, 5), None),   # one dead pair each way
             (None, None)]                                     # both dead: accept
    for lhs, rhs in cases:
        acc, gt = pf.decide(lhs, rhs, g2, s_g2)
        eacc, egt = m.kzg_decide(lhs, rhs, g2, s_g2)
        assert acc == eacc and gt == pf.w_from_model(egt)
