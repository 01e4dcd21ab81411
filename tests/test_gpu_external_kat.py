"""EXTERNAL known answers on the device: Ethereum's alt_bn128 precompile vectors (EIP-196 bn256Add / bn256ScalarMul, EIP-197
bn256Pairing; tests/golden/eip_vectors.json, validated by oracle/gen_golden_eip.py) through the C ABI, plus GT self-consistency
(bilinearity in the exponent) on the device's own GT bytes.  The reference ships no known-answer vectors for this path;
these are the anchors that do not come from this repository's own models."""
import pytest

import snark_verifier_b200 as sv
from eip_helpers import as_deciding_key, g1_bytes, h, pairing_operands
from oracle import bn254_model as m

pytestmark = pytest.mark.gpu
le = m.fe_to_le


@pytest.fixture(scope="module")
def loader():
    L = sv.CudaLoader(0)
    yield L
    L.close()


def test_eip196_scalar_mul_and_add_on_device(loader, golden):
    g = golden("eip_vectors")
    for v in g["scalar_mul"]:
        pt, exp = g1_bytes(v["x"], v["y"]), g1_bytes(v["out_x"], v["out_y"])
        s = le(h(v["scalar"]) % m.R)
        assert loader.msm(s, pt, 1) == exp, v["name"]                       # Pippenger pipeline, one term
        assert loader.msm_batch(s, pt, [0, 1])[0] == exp, v["name"]        # the literal per-term path (native.rs:61-71)
    # all six products in ONE MSM: sum of the public outputs (added with the model) — exercises bucket accumulation on KAT points
    s_all = b"".join(le(h(v["scalar"]) % m.R) for v in g["scalar_mul"])
    p_all = b"".join(g1_bytes(v["x"], v["y"]) for v in g["scalar_mul"])
    acc = None
    for v in g["scalar_mul"]:
        acc = m.g1_add(acc, (h(v["out_x"]), h(v["out_y"])))
    assert loader.msm(s_all, p_all, len(g["scalar_mul"])) == m.g1_to_bytes(acc)
    for v in g["add"]:
        a, b, exp = g1_bytes(v["x1"], v["y1"]), g1_bytes(v["x2"], v["y2"]), g1_bytes(v["out_x"], v["out_y"])
        assert loader.msm(le(1) * 2, a + b, 2) == exp, v["name"]


@pytest.mark.parametrize("mode", [1, 3, 4, 5], ids=["thread_per_check", "block_per_check", "warp_per_check", "latency_kernel"])
def test_eip197_pairing_vectors_on_device(golden, mode):
    """e(P1, Q1) e(P2, Q2) = 1 as a KZG decision with g2 = Q1, s_g2 = -Q2 (decider.rs:74-78); arbitrary G2 keys exercise
    k_g2_prepare on operands that are not multiples chosen by this repository.  Reject twin: P2 negated."""
    g = golden("eip_vectors")
    one = m.gt_to_bytes(m.f12_one())
    for v in g["pairing"]:
        L = sv.CudaLoader(0)
        try:
            L.set_pairing_mode(mode)
            p1, q1, p2, q2 = pairing_operands(v["words"])
            g2, s_g2 = as_deciding_key(q1, q2)
            kz = sv.KzgAs(L, sv.KzgDecidingKey(m.g1_to_bytes(m.G1_GEN), g2, s_g2))
            lhs = m.g1_to_bytes(p1) * 2
            rhs = m.g1_to_bytes(p2) + m.g1_to_bytes(m.g1_neg(p2))
            acc, gt = kz.decide_batch(lhs, rhs, 2, want_gt=True)
            assert acc == b"\x01\x00", v["name"]
            assert gt[:384] == one and gt[384:] != one, v["name"]
        finally:
            L.close()


def test_gt_bilinearity_on_device_bytes(loader, golden):
    """e(aG, bG2) = e(G, G2)^(ab): the device's GT bytes for (lhs = aG, rhs = O, g2 = bG2) against the golden e(G, G2) raised to ab
    with the Python model's Fq12 — a self-consistency KAT on GT values (the one output no external vector pins)."""
    g = golden("pairing")
    e_bytes = bytes.fromhex(g["e_G1_G2"])
    coeffs = [int.from_bytes(e_bytes[32 * i:32 * i + 32], "little") for i in range(12)]
    # tower order c0.c0.c0, c0.c0.c1, c0.c1.c0, ... -> w-power coefficients [c00, c10, c01, c11, c02, c12]
    tower = [(coeffs[2 * i], coeffs[2 * i + 1]) for i in range(6)]          # c00 c01 c02 c10 c11 c12
    e = [tower[0], tower[3], tower[1], tower[4], tower[2], tower[5]]
    assert m.gt_to_bytes(e) == e_bytes
    a, b = 0xC0FFEE1234567, 0xBADC0DE7654321
    bg2 = m.g2_to_bytes(m.g2_mul(m.G2_GEN, b))
    L = sv.CudaLoader(0)
    try:
        kz = sv.KzgAs(L, sv.KzgDecidingKey(m.g1_to_bytes(m.G1_GEN), bg2, bg2))
        _, gt = kz.decide_batch(m.g1_to_bytes(m.g1_mul(m.G1_GEN, a)), bytes(64), 1, want_gt=True)
    finally:
        L.close()
    assert gt == m.gt_to_bytes(m.f12_pow(e, a * b % m.R))
