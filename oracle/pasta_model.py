"""Independent Python big-integer model of the Pallas side of the path (TEST INFRASTRUCTURE ONLY): the curve, the reference's
Pippenger consumer `IpaAs::decide` (snark-verifier/src/pcs/ipa/decider.rs:47-70) and `h_coeffs` (pcs/ipa.rs:401-417).

Parity status: unpinned by the reference (it holds no known-answer vector for Pallas arithmetic; its IPA tests only accept).  What
pins this model: the Pasta parameters are public (Zcash Pasta curves: y^2 = x^3 + 5, generator (-1, 2), the 255-bit moduli below) and
self-checking — [q] G = O holds only if p, q and b are all right (asserted at import) — and every output is canonical mathematics
(an affine point), so any correct implementation produces the same bytes."""
P = 0x40000000000000000000000000000000224698FC094CF91B992D30ED00000001      # Pallas base field
Q = 0x40000000000000000000000000000000224698FC0994A8DD8C46EB2100000001      # Pallas scalar field = group order
B = 5
GEN = (P - 1, 2)
MASK64 = (1 << 64) - 1


def is_on_curve(pt):
    return pt is None or (pt[1] * pt[1] - pt[0] ** 3 - B) % P == 0


# Jacobian arithmetic (a = 0): fast enough for 10^3-term naive folds in pure Python
def _jdbl(X, Y, Z):
    if Y == 0 or Z == 0:
        return (0, 1, 0)
    A, Bq = X * X % P, Y * Y % P
    C = Bq * Bq % P
    D = 2 * ((X + Bq) * (X + Bq) - A - C) % P
    E = 3 * A % P
    X3 = (E * E - 2 * D) % P
    return (X3, (E * (D - X3) - 8 * C) % P, 2 * Y * Z % P)


def _jadd(a, b):
    X1, Y1, Z1 = a
    X2, Y2, Z2 = b
    if Z1 == 0:
        return b
    if Z2 == 0:
        return a
    Z1Z1, Z2Z2 = Z1 * Z1 % P, Z2 * Z2 % P
    U1, U2 = X1 * Z2Z2 % P, X2 * Z1Z1 % P
    S1, S2 = Y1 * Z2 * Z2Z2 % P, Y2 * Z1 * Z1Z1 % P
    if U1 == U2:
        return _jdbl(X1, Y1, Z1) if S1 == S2 else (0, 1, 0)
    H, Rr = (U2 - U1) % P, (S2 - S1) % P
    HH = H * H % P
    HHH = H * HH % P
    V = U1 * HH % P
    X3 = (Rr * Rr - HHH - 2 * V) % P
    return (X3, (Rr * (V - X3) - S1 * HHH) % P, Z1 * Z2 * H % P)


def _to_affine(j):
    if j[2] == 0:
        return None
    zi = pow(j[2], -1, P)
    return (j[0] * zi * zi % P, j[1] * zi * zi * zi % P)


def _from_affine(pt):
    return (0, 1, 0) if pt is None else (pt[0], pt[1], 1)


def add(a, b):
    return _to_affine(_jadd(_from_affine(a), _from_affine(b)))


def neg(pt):
    return None if pt is None else (pt[0], (-pt[1]) % P)


def _jmul(pt, k):
    acc, base = (0, 1, 0), _from_affine(pt)
    k %= Q
    while k:
        if k & 1:
            acc = _jadd(acc, base)
        base = _jdbl(*base)
        k >>= 1
    return acc


def mul(pt, k):
    return _to_affine(_jmul(pt, k))


def msm_naive(scalars, points):
    """loader/native.rs:61-71 semantics: fold of base * scalar, then to_affine"""
    acc = (0, 1, 0)
    for s, pt in zip(scalars, points):
        acc = _jadd(acc, _jmul(pt, s))
    return _to_affine(acc)


assert is_on_curve(GEN) and mul(GEN, Q) is None and mul(GEN, Q - 1) == neg(GEN)


def h_coeffs(xi, scalar=1):
    """pcs/ipa.rs:401-417, literally: the vector doubles once per challenge, taken in reverse order"""
    coeffs = [0] * (1 << len(xi))
    coeffs[0] = scalar % Q
    for i, x in enumerate(reversed(xi)):
        ln = 1 << i
        for j in range(ln):
            coeffs[ln + j] = coeffs[j] * x % Q
    return coeffs


def ipa_decide(g, u, xi):
    """pcs/ipa/decider.rs:47-59"""
    return u == msm_naive(h_coeffs(xi, 1), g)


# byte encodings (include/snarkv_cuda.h): canonical little-endian, identity = zeros
def fe_to_le(x):
    return int(x).to_bytes(32, "little")


def pt_to_bytes(pt):
    return bytes(64) if pt is None else fe_to_le(pt[0]) + fe_to_le(pt[1])


def pt_from_bytes(b):
    x, y = int.from_bytes(b[:32], "little"), int.from_bytes(b[32:64], "little")
    return None if x == 0 and y == 0 else (x, y)


# the deterministic generators of csrc/msm.cu (k_synth_scalars / k_synth_points), restated
def splitmix64(x):
    x = (x + 0x9E3779B97F4A7C15) & MASK64
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK64
    return z ^ (z >> 31)


def synth_scalar(seed, i):
    limbs = [splitmix64((seed * 0x100000001B3 + i * 4 + k) & MASK64) for k in range(4)]
    return limbs[0] | (limbs[1] << 64) | (limbs[2] << 128) | ((limbs[3] & ((1 << 62) - 1)) << 192)      # < 2^254 < q: already canonical


def synth_point_scalar(seed, i):
    return splitmix64((seed * 0x100000001B3 + 0x5151515151515151 + i) & MASK64) | 1
