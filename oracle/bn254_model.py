"""Independent Python big-integer model of the BN254 hot path (TEST INFRASTRUCTURE ONLY).

This file is the *second*, independent statement of the mathematics that the C++ oracle
(`oracle/oracle.cpp`) and the CUDA product path must agree with.  It deliberately uses different
algorithms from both of them so that agreement means something:

  * affine short-Weierstrass formulas with a modular inversion per operation (the oracle / CUDA path use
    Jacobian / XYZZ coordinates),
  * a plain binary (not NAF) Miller loop with affine line functions,
  * the final exponentiation as one naive `f ** ((p**12 - 1) // r)` square-and-multiply
    (the oracle / CUDA path use easy part + Frobenius maps + cyclotomic squarings).

What it models (reference = /root/reference/snark-verifier/src, arithmetic = halo2curves 0.6.0 `bn256`,
which is NOT vendored in the reference tree):

  * `NativeLoader::multi_scalar_multiplication` loader/native.rs:61-71   -> `msm_naive`
  * `KzgAs::decide` pcs/kzg/decider.rs:70-82                              -> `kzg_decide`
  * `KzgAs::verify` pcs/kzg/accumulation.rs:41-63                         -> `kzg_accumulate`

Parity status: **parity unpinned** by the reference (it ships no known-answer vectors); every output modelled
here is mathematically canonical (affine G1 point; GT = f^((p^12-1)/r) of the optimal-ate Miller value), so
this model + the C++ oracle agreeing bit-for-bit is the pin we can have without a Rust toolchain.

Only `tests/`, `oracle/gen_golden.py` and `__graft_entry__.smoke()` may import this module.
"""

# --------------------------------------------------------------------------------------------------
# Parameters (identities re-asserted below at import time)
# --------------------------------------------------------------------------------------------------
U = 4965661367192848881
P = 36 * U**4 + 36 * U**3 + 24 * U**2 + 6 * U + 1
R = 36 * U**4 + 36 * U**3 + 18 * U**2 + 6 * U + 1
ATE_LOOP = 6 * U + 2
assert P == 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
assert R == 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
assert P % 4 == 3 and P % 6 == 1

B1 = 3  # G1: y^2 = x^3 + 3
G1_GEN = (1, 2)

# --------------------------------------------------------------------------------------------------
# Fq2 = Fq[i]/(i^2+1), elements are (a, b) = a + b i
# --------------------------------------------------------------------------------------------------
def f2_add(x, y): return ((x[0] + y[0]) % P, (x[1] + y[1]) % P)
def f2_sub(x, y): return ((x[0] - y[0]) % P, (x[1] - y[1]) % P)
def f2_neg(x): return ((-x[0]) % P, (-x[1]) % P)
def f2_mul(x, y): return ((x[0] * y[0] - x[1] * y[1]) % P, (x[0] * y[1] + x[1] * y[0]) % P)
def f2_sqr(x): return f2_mul(x, x)
def f2_scalar(x, k): return (x[0] * k % P, x[1] * k % P)
def f2_conj(x): return (x[0], (-x[1]) % P)
def f2_inv(x):
    n = pow((x[0] * x[0] + x[1] * x[1]) % P, -1, P)
    return (x[0] * n % P, (-x[1]) * n % P)
def f2_pow(x, e):
    r = (1, 0)
    while e:
        if e & 1: r = f2_mul(r, x)
        x = f2_sqr(x); e >>= 1
    return r
F2_ZERO, F2_ONE = (0, 0), (1, 0)
XI = (9, 1)  # non-residue for Fq6 = Fq2[v]/(v^3 - xi)

# --------------------------------------------------------------------------------------------------
# Fq12 modelled DIRECTLY as Fq2[w]/(w^6 - xi): a list of 6 Fq2 coefficients of w^0..w^5.
# (The oracle / CUDA path use the 2-3-2 tower Fq12 = Fq6[w]/(w^2 - v), Fq6 = Fq2[v]/(v^3 - xi); with w^2 = v the
#  tower element c0 + c1 w, c_k = c_k0 + c_k1 v + c_k2 v^2 has w-power coefficients
#  [c00, c10, c01, c11, c02, c12].  `f12_to_tower` converts.)
# --------------------------------------------------------------------------------------------------
def f12_one(): return [F2_ONE] + [F2_ZERO] * 5
def f12_mul(x, y):
    t = [F2_ZERO] * 11
    for i in range(6):
        if x[i] == F2_ZERO: continue
        for j in range(6):
            if y[j] == F2_ZERO: continue
            t[i + j] = f2_add(t[i + j], f2_mul(x[i], y[j]))
    for k in range(10, 5, -1):
        t[k - 6] = f2_add(t[k - 6], f2_mul(t[k], XI))
    return t[:6]
def f12_pow(x, e):
    r = f12_one()
    for bit in bin(e)[2:]:
        r = f12_mul(r, r)
        if bit == '1': r = f12_mul(r, x)
    return r
def f12_to_tower(x):
    """-> 12 Fq ints in the serialisation order fixed by include/snarkv_cuda.h:
    c0.c0.c0, c0.c0.c1, c0.c1.c0, c0.c1.c1, c0.c2.c0, c0.c2.c1, c1.c0.c0, ... c1.c2.c1"""
    order = [x[0], x[2], x[4], x[1], x[3], x[5]]
    out = []
    for c in order: out += [c[0], c[1]]
    return out

# --------------------------------------------------------------------------------------------------
# G1 (affine, None = identity)
# --------------------------------------------------------------------------------------------------
def g1_is_on_curve(pt):
    if pt is None: return True
    x, y = pt
    return (y * y - x * x * x - B1) % P == 0
def g1_neg(pt): return None if pt is None else (pt[0], (-pt[1]) % P)
def g1_add(a, b):
    if a is None: return b
    if b is None: return a
    x1, y1 = a; x2, y2 = b
    if x1 == x2:
        if (y1 + y2) % P == 0: return None
        lam = 3 * x1 * x1 * pow(2 * y1, -1, P) % P
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, P) % P
    x3 = (lam * lam - x1 - x2) % P
    return (x3, (lam * (x1 - x3) - y1) % P)
def g1_mul(pt, k):
    k %= R
    acc = None
    while k:
        if k & 1: acc = g1_add(acc, pt)
        pt = g1_add(pt, pt); k >>= 1
    return acc
def msm_naive(scalars, points):
    """loader/native.rs:61-71: fold of base*scalar, then to_affine (canonical affine result)."""
    assert len(scalars) == len(points) and len(scalars) > 0  # .unwrap() on empty, native.rs:69
    acc = None
    for s, pt in zip(scalars, points):
        acc = g1_add(acc, g1_mul(pt, s))
    return acc

# --------------------------------------------------------------------------------------------------
# G2 on the D-type sextic twist  y^2 = x^3 + 3/xi  over Fq2 (affine, None = identity)
# --------------------------------------------------------------------------------------------------
B2 = f2_mul((3, 0), f2_inv(XI))
assert B2 == (0x2B149D40CEB8AAAE81BE18991BE06AC3B5B4C5E559DBEFA33267E6DC24A138E5,
              0x009713B03AF0FED4CD2CAFADEED8FDF4A74FA084E52D1852E4A2BD0685C315D2)
G2_GEN = ((10857046999023057135944570762232829481370756359578518086990519993285655852781,
           11559732032986387107991004021392285783925812861821192530917403151452391805634),
          (8495653923123431417604973247489272438418190587263600148770280649306958101930,
           4082367875863433681332203403145435568316851327593401208105741076214120093531))
def g2_is_on_curve(pt):
    if pt is None: return True
    x, y = pt
    return f2_sub(f2_sqr(y), f2_add(f2_mul(f2_sqr(x), x), B2)) == F2_ZERO
def g2_neg(pt): return None if pt is None else (pt[0], f2_neg(pt[1]))
def g2_add(a, b):
    if a is None: return b
    if b is None: return a
    x1, y1 = a; x2, y2 = b
    if x1 == x2:
        if f2_add(y1, y2) == F2_ZERO: return None
        lam = f2_mul(f2_scalar(f2_sqr(x1), 3), f2_inv(f2_scalar(y1, 2)))
    else:
        lam = f2_mul(f2_sub(y2, y1), f2_inv(f2_sub(x2, x1)))
    x3 = f2_sub(f2_sub(f2_sqr(lam), x1), x2)
    return (x3, f2_sub(f2_mul(lam, f2_sub(x1, x3)), y1))
def g2_mul(pt, k):
    k %= R
    acc = None
    while k:
        if k & 1: acc = g2_add(acc, pt)
        pt = g2_add(pt, pt); k >>= 1
    return acc
assert g2_is_on_curve(G2_GEN)

# --------------------------------------------------------------------------------------------------
# Optimal-ate pairing, affine lines, binary loop.
#   untwist psi(x', y') = (x' w^2, y' w^3); the line through psi(T) with twist-slope lam evaluated at P=(xP,yP):
#     l = yP  -  lam * xP * w  +  (lam * xT - yT) * w^3          (vertical lines dropped: they lie in Fq6)
# --------------------------------------------------------------------------------------------------
def _line(T, lam, Pt):
    xP, yP = Pt
    l = [F2_ZERO] * 6
    l[0] = (yP % P, 0)
    l[1] = f2_neg(f2_scalar(lam, xP))
    l[3] = f2_sub(f2_mul(lam, T[0]), T[1])
    return l
def _dbl_line(T, Pt):
    lam = f2_mul(f2_scalar(f2_sqr(T[0]), 3), f2_inv(f2_scalar(T[1], 2)))
    return _line(T, lam, Pt), g2_add(T, T)
def _add_line(T, Q, Pt):
    lam = f2_mul(f2_sub(Q[1], T[1]), f2_inv(f2_sub(Q[0], T[0])))
    return _line(T, lam, Pt), g2_add(T, Q)
GAMMA12 = f2_pow(XI, (P - 1) // 3)
GAMMA13 = f2_pow(XI, (P - 1) // 2)
def g2_frobenius(Q):
    """pi on the twist: psi^-1 . Frob_p . psi"""
    return (f2_mul(f2_conj(Q[0]), GAMMA12), f2_mul(f2_conj(Q[1]), GAMMA13))
def miller_loop(pairs):
    """Product of optimal-ate Miller functions f_{6u+2,Q}(P) * l_{[6u+2]Q, pi(Q)}(P) * l_{.., -pi^2(Q)}(P).
    Pairs with an identity on either side are skipped (they contribute 1), as halo2curves' multi_miller_loop does."""
    f = f12_one()
    for Pt, Q in pairs:
        if Pt is None or Q is None: continue
        T = Q
        g = f12_one()
        for bit in bin(ATE_LOOP)[3:]:
            l, T = _dbl_line(T, Pt)
            g = f12_mul(f12_mul(g, g), l)
            if bit == '1':
                l, T = _add_line(T, Q, Pt)
                g = f12_mul(g, l)
        Q1 = g2_frobenius(Q)
        Q2 = g2_neg(g2_frobenius(Q1))
        l, T = _add_line(T, Q1, Pt); g = f12_mul(g, l)
        l, T = _add_line(T, Q2, Pt); g = f12_mul(g, l)
        f = f12_mul(f, g)
    return f
FINAL_EXP = (P**12 - 1) // R
def final_exponentiation(f): return f12_pow(f, FINAL_EXP)
def pairing(Pt, Q): return final_exponentiation(miller_loop([(Pt, Q)]))

def kzg_decide(lhs, rhs, g2, s_g2):
    """pcs/kzg/decider.rs:70-82: accept iff e(lhs, g2) * e(rhs, -s_g2) == 1.  Returns (accept, gt)."""
    gt = final_exponentiation(miller_loop([(lhs, g2), (rhs, g2_neg(s_g2))]))
    return gt == f12_one(), gt

def kzg_accumulate(accs, r, blind=None):
    """pcs/kzg/accumulation.rs:41-63: (sum r^i lhs_i, sum r^i rhs_i), blind pair chained last."""
    pairs = list(accs) + ([blind] if blind is not None else [])
    powers = [pow(r, i, R) for i in range(len(pairs))]   # loader.rs:71-78 `powers`
    return (msm_naive(powers, [a[0] for a in pairs]), msm_naive(powers, [a[1] for a in pairs]))

# --------------------------------------------------------------------------------------------------
# Byte encodings shared with include/snarkv_cuda.h
# --------------------------------------------------------------------------------------------------
def fe_to_le(x): return int(x).to_bytes(32, 'little')
def g1_to_bytes(pt):
    """x||y canonical little-endian, identity = 64 zero bytes (halo2curves encodes identity as (0,0))."""
    if pt is None: return bytes(64)
    return fe_to_le(pt[0]) + fe_to_le(pt[1])
def g1_from_bytes(b):
    x = int.from_bytes(b[:32], 'little'); y = int.from_bytes(b[32:64], 'little')
    return None if x == 0 and y == 0 else (x, y)
def g2_to_bytes(pt):
    """x.c0||x.c1||y.c0||y.c1 canonical LE, identity = 128 zero bytes."""
    if pt is None: return bytes(128)
    return b''.join(fe_to_le(c) for c in (pt[0][0], pt[0][1], pt[1][0], pt[1][1]))
def gt_to_bytes(f): return b''.join(fe_to_le(c) for c in f12_to_tower(f))

# deterministic test-data generator shared (by definition) with oracle.cpp and the CUDA generator kernel
MASK64 = (1 << 64) - 1
def splitmix64(x):
    x = (x + 0x9E3779B97F4A7C15) & MASK64
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK64
    return z ^ (z >> 31)
def synth_scalar(seed, i):
    """4 x splitmix64 limbs (LE), top limb masked to 62 bits, then one conditional subtract of r  => uniform-ish in [0,r)."""
    limbs = [splitmix64((seed * 0x100000001B3 + i * 4 + k) & MASK64) for k in range(4)]
    v = limbs[0] | (limbs[1] << 64) | (limbs[2] << 128) | ((limbs[3] & ((1 << 62) - 1)) << 192)
    return v - R if v >= R else v
def synth_point_scalar(seed, i):
    """64-bit multiplier t_i with P_i = [t_i]G (never 0)."""
    return splitmix64((seed * 0x100000001B3 + 0x5151515151515151 + i) & MASK64) | 1
