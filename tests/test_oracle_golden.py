"""The C++ oracle (oracle/oracle.cpp) against the golden vectors produced by the independent Python big-int model
(oracle/gen_golden.py).  CPU only.  This is the pin that replaces the reference's missing known-answer vectors
(SURVEY.md §8c): two independent implementations agreeing bit-for-bit on canonical outputs."""
import pytest

import oracle
from oracle import bn254_model as m

H = bytes.fromhex


def test_field_constants_match_survey_appendix(golden):
    g = golden("field")
    # Montgomery constants quoted in SURVEY.md appendix A (halo2curves layout: 4 x u64 LE limbs)
    assert g["fq"]["inv64"] == "87d20782e4866389"
    assert g["fr"]["inv64"] == "c2e1f593efffffff"
    assert int.from_bytes(H(g["fq"]["mont_R"]), "little") == 0x0e0a77c19a07df2f666ea36f7879462c0a78eb28f5c70b3dd35d438dc58f0d9d
    assert int.from_bytes(H(g["fr"]["mont_R"]), "little") == 0x0e0a77c19a07df2f666ea36f7879462e36fc76959f60cd29ac96341c4ffffffb
    # oracle's Montgomery form of 1 is R
    one = (1).to_bytes(32, "little")
    assert oracle.fp_to_mont(0, one) == H(g["fq"]["mont_R"])
    assert oracle.fp_to_mont(1, one) == H(g["fr"]["mont_R"])


@pytest.mark.parametrize("field,name", [(0, "fq"), (1, "fr")])
def test_field_mul_inv(golden, field, name):
    g = golden("field")[name]
    for a, b, c in g["mul"]:
        assert oracle.fp_mul(field, H(a), H(b)) == H(c)
    for a, ai in g["inv"]:
        assert oracle.fp_inv(field, H(a)) == H(ai)


def test_field_rejects_non_canonical(golden):
    p = H(golden("field")["p"])
    with pytest.raises(ValueError):
        oracle.fp_mul(0, p, p)  # from_repr fails for values >= modulus


def test_g1_scalar_mul_and_add(golden):
    g = golden("g1")
    gen = H(g["generator"])
    for k, exp in g["mul_G"]:
        assert oracle.g1_mul(gen, H(k)) == H(exp)
    for a, b, c in g["add"]:
        assert oracle.g1_add(H(a), H(b)) == H(c)


def test_g1_rejects_off_curve():
    bad = (1).to_bytes(32, "little") + (3).to_bytes(32, "little")
    assert not oracle.g1_is_on_curve(bad)
    with pytest.raises(ValueError):
        oracle.g1_mul(bad, (5).to_bytes(32, "little"))


def test_msm_native_fold_and_pippenger(golden):
    for case in golden("msm"):
        s, p, n = H(case["scalars"]), H(case["points"]), case["n"]
        exp = H(case["expected"])
        assert oracle.msm_native(s, p, n) == exp, case["name"]            # loader/native.rs:61-71
        assert oracle.msm_pippenger(s, p, n, 1) == exp, case["name"]       # util/msm.rs:259-304
        assert oracle.msm_pippenger(s, p, n, 4) == exp, case["name"]       # util/msm.rs:308-343


def test_msm_raw_layout_entry_matches_canonical_entry():
    n = 500
    s = oracle.synth_scalars(17, 0, n); p = oracle.synth_points(17, 0, n, 4)
    sm = oracle.to_mont_batch(1, s, n, 3); pm = oracle.to_mont_batch(0, p, 2 * n, 3)
    assert sm[:32] == oracle.fp_to_mont(1, s[:32]) and pm[32:64] == oracle.fp_to_mont(0, p[32:64])
    assert oracle.msm_pippenger_raw(sm, pm, n, 4) == oracle.msm_pippenger(s, p, n, 4)


def test_msm_empty_is_an_error():
    with pytest.raises(ValueError):       # .unwrap() on an empty fold panics, native.rs:69
        oracle.msm_native(b"", b"", 0)


def test_synth_generators_match_python_model():
    n = 9
    assert oracle.synth_scalars(3, 5, n) == b"".join(m.fe_to_le(m.synth_scalar(3, 5 + i)) for i in range(n))
    ts = [m.synth_point_scalar(3, 5 + i) for i in range(n)]
    assert list(oracle.synth_point_scalars(3, 5, n)) == ts
    assert oracle.synth_points(3, 5, n, 2) == b"".join(m.g1_to_bytes(m.g1_mul(m.G1_GEN, t)) for t in ts)


def test_msm_dlog_checksum_agrees_with_pippenger():
    n = 3000
    s = oracle.synth_scalars(11, 0, n)
    p = oracle.synth_points(11, 0, n, 4)
    t = oracle.synth_point_scalars(11, 0, n)
    assert oracle.msm_pippenger(s, p, n, 4) == oracle.msm_expected_from_dlogs(s, t, n)


def test_pairing_gt_bytes_and_decide(golden):
    g = golden("pairing")
    g2, s_g2 = H(g["g2_generator"]), H(g["s_g2"])
    assert oracle.g2_generator() == g2
    assert oracle.g2_mul(g2, H(g["s"])) == s_g2
    # e(G1, G2): decide(lhs=G1, rhs=identity) leaves exactly e(G1,G2) in GT
    gen = m.g1_to_bytes(m.G1_GEN)
    ok, gt = oracle.kzg_decide(gen, bytes(64), g2, s_g2)
    assert not ok and gt == H(g["e_G1_G2"])
    for c in g["checks"]:
        ok, gt = oracle.kzg_decide(H(c["lhs"]), H(c["rhs"]), g2, s_g2)
        assert ok == c["accept"], c["name"]
        assert gt == H(c["gt"]), c["name"]


def test_decide_batch_matches_single_and_hoisted(golden):
    g = golden("pairing")
    g2, s_g2 = H(g["g2_generator"]), H(g["s_g2"])
    lhs = b"".join(H(c["lhs"]) for c in g["checks"])
    rhs = b"".join(H(c["rhs"]) for c in g["checks"])
    n = len(g["checks"])
    for hoist in (0, 1):
        acc, gt = oracle.kzg_decide_batch(lhs, rhs, n, g2, s_g2, threads=3, hoist=hoist, want_gt=True)
        assert list(acc) == [int(c["accept"]) for c in g["checks"]]
        assert gt == b"".join(H(c["gt"]) for c in g["checks"])


def test_pairing_bilinearity_property():
    # e(aP, bQ) e(-abP, Q) = 1   <=>  decide(lhs = a*b*G, rhs = a*G, g2, s_g2 = b*G2) accepts
    a, b = 0x1234567890ABCDEF1234567890ABCDEF, 0xFEDCBA0987654321
    g2 = oracle.g2_generator()
    bg2 = oracle.g2_mul(g2, m.fe_to_le(b))
    gen = m.g1_to_bytes(m.G1_GEN)
    lhs = oracle.g1_mul(gen, m.fe_to_le(a * b % m.R))
    rhs = oracle.g1_mul(gen, m.fe_to_le(a))
    assert oracle.kzg_decide(lhs, rhs, g2, bg2, want_gt=False)[0]
    assert not oracle.kzg_decide(lhs, oracle.g1_mul(gen, m.fe_to_le(a + 1)), g2, bg2, want_gt=False)[0]


def test_accumulate(golden):
    for c in golden("accumulate"):
        a, b = oracle.kzg_accumulate(H(c["lhs"]), H(c["rhs"]), c["n"], H(c["r"]))
        assert a == H(c["out_lhs"]) and b == H(c["out_rhs"])


def test_external_known_answers_from_ethereum_precompile_vectors():
    """The reference ships no known-answer vectors for this path, but BN254 is Ethereum's alt_bn128: the public
    EIP-196 / EIP-197 precompile test vectors (go-ethereum's bn256Add / bn256Pairing cases) pin the group law of both groups.
      * G1: (1, 2) + (1, 2) = 2 G  (EIP-196 `bn256Add` vector)
      * G2: 2 * G2_generator       (the G2 operand of the `bn256Pairing` two-point cases; EVM words are imaginary part first)
    Both the Python model and the C++ restatement must reproduce them bit for bit."""
    two_g1 = (0x030644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD3,
              0x15ED738C0E0A7C92E7845F96B2AE9C0A68A6A449E3538FC7FF3EBF7A5A18A2C4)
    two_g2_evm_words = (0x203E205DB4F19B37B60121B83A7333706DB86431C6D835849957ED8C3928AD79,    # x.c1
                        0x27DC7234FD11D3E8C36C59277C3E6F149D5CD3CFA9A62AEE49F8130962B4B3B9,    # x.c0
                        0x195E8AA5B7827463722B8C153931579D3505566B4EDF48D498E185F0509DE152,    # y.c1
                        0x04BB53B8977E5F92A0BC372742C4830944A59B4FE6B1C0466E2A6DAD122B5D2E)    # y.c0
    le = m.fe_to_le
    assert m.g1_mul(m.G1_GEN, 2) == two_g1 == m.g1_add(m.G1_GEN, m.G1_GEN)
    gen = m.g1_to_bytes(m.G1_GEN)
    assert oracle.g1_mul(gen, le(2)) == le(two_g1[0]) + le(two_g1[1])
    assert oracle.msm_native(le(1) * 2, gen * 2, 2) == le(two_g1[0]) + le(two_g1[1])
    x, y = m.g2_mul(m.G2_GEN, 2)
    assert (x[1], x[0], y[1], y[0]) == two_g2_evm_words
    x1, x0, y1, y0 = two_g2_evm_words
    assert oracle.g2_mul(oracle.g2_generator(), le(2)) == le(x0) + le(x1) + le(y0) + le(y1)


def test_eip196_eip197_public_vectors_pin_the_oracle(golden):
    """EXTERNAL known answers: Ethereum's alt_bn128 precompile vectors (tests/golden/eip_vectors.json, validated by
    oracle/gen_golden_eip.py) through the C++ restatement — scalar multiplication, addition, and three `bn256Pairing` accept cases
    run as KZG decisions (g2 = Q1, s_g2 = -Q2) plus their public reject twins (second G1 operand negated)."""
    from eip_helpers import as_deciding_key, g1_bytes, h, pairing_operands
    g = golden("eip_vectors")
    le = m.fe_to_le
    for v in g["scalar_mul"]:
        pt, exp = g1_bytes(v["x"], v["y"]), g1_bytes(v["out_x"], v["out_y"])
        s = le(h(v["scalar"]) % m.R)
        assert oracle.g1_mul(pt, s) == exp, v["name"]
        assert oracle.msm_native(s, pt, 1) == exp, v["name"]
        assert oracle.msm_pippenger(s, pt, 1) == exp, v["name"]
    for v in g["add"]:
        a, b, exp = g1_bytes(v["x1"], v["y1"]), g1_bytes(v["x2"], v["y2"]), g1_bytes(v["out_x"], v["out_y"])
        assert oracle.g1_add(a, b) == exp, v["name"]
        assert oracle.msm_native(le(1) * 2, a + b, 2) == exp, v["name"]
    for v in g["pairing"]:
        p1, q1, p2, q2 = pairing_operands(v["words"])
        g2, s_g2 = as_deciding_key(q1, q2)
        ok, gt = oracle.kzg_decide(m.g1_to_bytes(p1), m.g1_to_bytes(p2), g2, s_g2)
        assert ok and gt == m.gt_to_bytes(m.f12_one()), v["name"]
        ok, _ = oracle.kzg_decide(m.g1_to_bytes(p1), m.g1_to_bytes(m.g1_neg(p2)), g2, s_g2)
        assert not ok, v["name"]
