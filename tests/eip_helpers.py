"""Shared decoding of tests/golden/eip_vectors.json (public EIP-196 / EIP-197 precompile vectors) into this repo's byte formats."""
from oracle import bn254_model as m

le = m.fe_to_le


def h(s):
    return int(s, 16)


def g1_bytes(x, y):
    return le(h(x)) + le(h(y))


def pairing_operands(words):
    """12 EVM words (two (G1, G2) pairs, G2 coordinates imaginary part first) -> (P1, Q1, P2, Q2) as model points."""
    w = [h(t) for t in words]
    pairs = []
    for i in (0, 6):
        pairs.append(((w[i], w[i + 1]), ((w[i + 3], w[i + 2]), (w[i + 5], w[i + 4]))))
    return pairs[0][0], pairs[0][1], pairs[1][0], pairs[1][1]


def as_deciding_key(q1, q2):
    """e(P1, Q1) e(P2, Q2) as a KZG decision: decide computes e(lhs, g2) e(rhs, -s_g2) (decider.rs:74-78), so g2 = Q1, s_g2 = -Q2."""
    return m.g2_to_bytes(q1), m.g2_to_bytes(m.g2_neg(q2))
