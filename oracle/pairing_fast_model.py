"""Test infrastructure: a thread-by-thread Python emulation of the DATA FLOW of csrc/pairing_fast.cu (the latency-organised
block-per-check KZG decision) over exact integers, checked against the independent pairing of oracle/bn254_model.py.

What it pins (tests/test_pairing_fast_model.py): the term map of the 192-thread Fq12 product (16-lane groups, xi folded into
the operand, xor-butterfly sums), the Frobenius sign rule, the merged two-pair line tables (slots, per-check scalars, dead
pairs), the signed-digit exponentiation by u, the norm-based Fq12 inversion and the final-exponentiation chain — i.e.
everything in that kernel that is index arithmetic rather than field arithmetic.  The kernel uses the same slot numbers and
the same formulas; the field operations themselves are the library's generated PTX (verified separately).

Fq12 values are lists of 12 integers: word 2 i + part = coefficient of w^i (part 0 = real, 1 = imaginary), w^6 = xi = 9 + u.
"""
from . import bn254_model as m

P = m.P
NWORDS = 12


def w_from_model(x):
    """bn254_model Fq12 (6 Fq2 coefficients of w^0..w^5) -> 12 words"""
    out = []
    for c in x:
        out += [c[0] % P, c[1] % P]
    return out


def w_to_model(a):
    return [(a[2 * i], a[2 * i + 1]) for i in range(6)]


# ---- the 192-thread product -------------------------------------------------------------------------------------------
def term_map(t):
    """thread t < 192 -> (o, s, active, ia, jb, jo, high, c, part, odd): identical arithmetic to fast::TermMap"""
    o, s = t >> 4, t & 15
    k, part = o >> 1, o & 1
    active = s < 12
    i, odd = s >> 1, s & 1
    jj = k - i
    high = jj < 0
    j = jj + 6 if high else jj
    c = part ^ odd
    return o, s, active, 2 * i + odd, 2 * j + c, 2 * j + (c ^ 1), high, c, part, odd


def xi_words(b):
    """xi * b word by word: (9 re - im, 9 im + re)"""
    out = [0] * 12
    for j in range(6):
        out[2 * j] = (9 * b[2 * j] - b[2 * j + 1]) % P
        out[2 * j + 1] = (9 * b[2 * j + 1] + b[2 * j]) % P
    return out


def op_mul(a, b, bx=None):
    v = [0] * 192
    for t in range(192):
        o, s, active, ia, jb, jo, high, c, part, odd = term_map(t)
        if not active:
            continue
        if not high:
            x = b[jb]
        elif bx is not None:
            x = bx[jb]
        else:
            x = (9 * b[jb] + b[jo]) % P if c else (9 * b[jb] - b[jo]) % P
        v[t] = a[ia] * x % P
    # level 1: re words subtract (even - odd), im words add; then xor-butterfly 2, 4, 8 inside every 16-lane group
    r = [0] * 192
    for t in range(192):
        o, s, active, ia, jb, jo, high, c, part, odd = term_map(t)
        pv = v[t ^ 1]
        if part:
            r[t] = (v[t] + pv) % P
        else:
            r[t] = (pv - v[t]) % P if odd else (v[t] - pv) % P
    for d in (2, 4, 8):
        r = [(r[t] + r[t ^ d]) % P for t in range(192)]
    return [r[16 * o] for o in range(12)]


def gamma(k):
    g1 = m.f2_pow(m.XI, (P - 1) // 6)
    g2 = m.f2_mul(g1, m.f2_conj(g1))
    g3 = m.f2_mul(g1, g2)
    g = {1: g1, 2: g2, 3: g3}[k]
    out, pw = [], (1, 0)
    for _ in range(6):
        out.append(pw)
        pw = m.f2_mul(pw, g)
    return out


def op_frobenius(a, k):
    """word (i, part): lanes s = 0, 1 of its group multiply a_re / a_im by a gamma part; one combine level"""
    g = gamma(k)
    cj = k & 1
    out = [0] * 12
    for o in range(12):
        i, part = o >> 1, o & 1
        v0 = a[2 * i] * (g[i][1] if part else g[i][0]) % P
        v1 = a[2 * i + 1] * (g[i][0] if part else g[i][1]) % P
        sub = (part ^ cj) == 0
        out[o] = (v0 - v1) % P if sub else (v0 + v1) % P
    return out


def op_conj(a):
    return [(-a[o]) % P if ((o >> 1) & 1) else a[o] for o in range(12)]


# ---- merged lines -------------------------------------------------------------------------------------------------------
def ate_naf():
    n, naf = m.ATE_LOOP, []
    while n:
        d = 0
        if n & 1:
            d = 2 - (n & 3)
            n -= d
        naf.append(d)
        n >>= 1
    return naf


def line_coeffs(Q):
    """(cy, cx, c0) per step with l(P) = cy yP + cx xP w + c0 w^3 — AFFINE lines here (the device's are projectively scaled by
    Fq2 factors that the final exponentiation kills; the slot algebra is the same)."""
    naf = ate_naf()
    T = Q
    out = []

    def dbl(T):
        lam = m.f2_mul(m.f2_scalar(m.f2_sqr(T[0]), 3), m.f2_inv(m.f2_scalar(T[1], 2)))
        return ((1, 0), m.f2_neg(lam), m.f2_sub(m.f2_mul(lam, T[0]), T[1])), m.g2_add(T, T)

    def add(T, S):
        lam = m.f2_mul(m.f2_sub(S[1], T[1]), m.f2_inv(m.f2_sub(S[0], T[0])))
        return ((1, 0), m.f2_neg(lam), m.f2_sub(m.f2_mul(lam, T[0]), T[1])), m.g2_add(T, S)

    nQ = m.g2_neg(Q)
    for i in range(len(naf) - 2, -1, -1):
        l, T = dbl(T)
        out.append(l)
        if naf[i] == 1:
            l, T = add(T, Q)
            out.append(l)
        elif naf[i] == -1:
            l, T = add(T, nQ)
            out.append(l)
    Q1 = m.g2_frobenius(Q)
    Q2 = m.g2_neg(m.g2_frobenius(Q1))
    l, T = add(T, Q1)
    out.append(l)
    l, T = add(T, Q2)
    out.append(l)
    return out


NSLOT = 16


def pair_tables(co0, co1):
    """per step 16 Fq2 slots for the three modes (both pairs live / only pair 0 / only pair 1) — fast::k_pair_tables"""
    z = (0, 0)
    both, only0, only1 = [], [], []
    for (cy1, cx1, c01), (cy2, cx2, c02) in zip(co0, co1):
        s = [z] * NSLOT
        s[0] = m.f2_mul(cy1, cy2)
        s[1] = m.f2_mul(m.XI, m.f2_mul(c01, c02))
        s[2] = m.f2_mul(cy1, cx2)
        s[3] = m.f2_mul(cx1, cy2)
        s[4] = m.f2_mul(cx1, cx2)
        s[6] = m.f2_mul(cy1, c02)
        s[7] = m.f2_mul(c01, cy2)
        s[8] = m.f2_mul(cx1, c02)
        s[9] = m.f2_mul(c01, cx2)
        both.append(s)
        for dst, (cy, cx, c0) in ((only0, (cy1, cx1, c01)), (only1, (cy2, cx2, c02))):
            s = [z] * NSLOT
            s[0], s[2], s[6] = cy, cx, c0
            dst.append(s)
    return both, only0, only1


def check_scalars(p0, p1, live0, live1):
    """the 16 per-check Fq scalars S[slot]"""
    S = [0] * NSLOT
    if live0 and live1:
        (x1, y1), (x2, y2) = p0, p1
        S[0], S[1], S[2], S[3], S[4] = y1 * y2 % P, 1, y1 * x2 % P, x1 * y2 % P, x1 * x2 % P
        S[6], S[7], S[8], S[9] = y1, y2, x1, x2
    else:
        x, y = p0 if live0 else p1
        S[0], S[2], S[6] = y, x, 1
    return S


def merged_line(slots, S):
    """one warp per step: lane = 2 slot + part computes slot.part * S[slot]; partner lane xor 2 added; even slots 0..10 hold
    L_0..L_5 (L_5 = 0); returns (L words, xi L words)"""
    v = [0] * 32
    for lane in range(32):
        slot, part = lane >> 1, lane & 1
        v[lane] = slots[slot][part] * S[slot] % P
    r = [(v[lane] + v[lane ^ 2]) % P for lane in range(32)]
    L = [0] * 12
    for lane in range(32):
        slot, part = lane >> 1, lane & 1
        if (slot & 1) == 0 and slot < 12:
            L[2 * (slot >> 1) + part] = r[lane]
    return L, xi_words(L)


# ---- exponentiation by u in signed digits -------------------------------------------------------------------------------
def u_naf_msb_first():
    n, naf = m.U, []
    while n:
        d = 0
        if n & 1:
            d = 2 - (n & 3)
            n -= d
        naf.append(d)
        n >>= 1
    assert sum(d << i for i, d in enumerate(naf)) == m.U and naf[-1] == 1
    return naf[::-1]


def op_exp_by_u(a):
    """a in the cyclotomic subgroup (inverse = conjugate)"""
    base, basex = a, xi_words(a)
    basec, basecx = op_conj(base), op_conj(basex)
    digits = u_naf_msb_first()
    r = list(a)
    for d in digits[1:]:
        r = op_mul(r, r)
        if d == 1:
            r = op_mul(r, base, basex)
        elif d == -1:
            r = op_mul(r, basec, basecx)
    return r


# ---- inversion by norms -------------------------------------------------------------------------------------------------
def op_inverse(a):
    c = op_conj(a)                       # a^(p^6)
    t = op_mul(a, c)                     # in Fq6 (even powers of w)
    t2 = op_frobenius(t, 2)
    t4 = op_frobenius(t2, 2)
    s = op_mul(t2, t4)
    n = op_mul(t, s)                     # in Fq2: words 0, 1
    assert all(x == 0 for x in n[2:]), "norm did not land in Fq2"
    d = pow((n[0] * n[0] + n[1] * n[1]) % P, -1, P)
    ninv = [n[0] * d % P, (-n[1]) * d % P] + [0] * 10
    sinv = op_mul(s, ninv)               # t^-1
    return op_mul(c, sinv)


def final_exponentiation(f):
    finv = op_inverse(f)
    f = op_mul(op_conj(f), finv)
    f = op_mul(op_frobenius(f, 2), f)
    fu = op_exp_by_u(f)
    fu2 = op_exp_by_u(fu)
    fu3 = op_exp_by_u(fu2)
    y0 = op_mul(op_mul(op_frobenius(f, 1), op_frobenius(f, 2)), op_frobenius(f, 3))
    y1 = op_conj(f)
    y2 = op_frobenius(fu2, 2)
    y3 = op_conj(op_frobenius(fu, 1))
    y4 = op_conj(op_mul(op_frobenius(fu2, 1), fu))
    y5 = op_conj(fu2)
    y6 = op_conj(op_mul(op_frobenius(fu3, 1), fu3))
    t0 = op_mul(op_mul(op_mul(y6, y6), y4), y5)
    t1 = op_mul(op_mul(y3, y5), t0)
    t0 = op_mul(t0, y2)
    t1 = op_mul(t1, t1)
    t1 = op_mul(t1, t0)
    t1 = op_mul(t1, t1)
    t0 = op_mul(t1, y1)
    t1 = op_mul(t1, y0)
    t0 = op_mul(t0, t0)
    return op_mul(t0, t1)


def decide(lhs, rhs, g2, s_g2):
    """the kernel's whole schedule for one check; returns (accept, gt words)"""
    Q0, Q1 = g2, m.g2_neg(s_g2)
    live0 = lhs is not None and Q0 is not None
    live1 = rhs is not None and Q1 is not None
    f = [1] + [0] * 11
    if live0 or live1:
        co0 = line_coeffs(Q0) if Q0 is not None else None
        co1 = line_coeffs(Q1) if Q1 is not None else None
        z3 = ((0, 0), (0, 0), (0, 0))
        both, only0, only1 = pair_tables(co0 or [z3] * len(co1), co1 or [z3] * len(co0))
        table = both if (live0 and live1) else (only0 if live0 else only1)
        S = check_scalars(lhs, rhs, live0, live1)
        lines = [merged_line(slots, S) for slots in table]
        naf = ate_naf()
        idx = 0
        for b in range(len(naf) - 2, -1, -1):
            if b != len(naf) - 2:
                f = op_mul(f, f)
            f = op_mul(f, lines[idx][0], lines[idx][1]); idx += 1
            if naf[b] != 0:
                f = op_mul(f, lines[idx][0], lines[idx][1]); idx += 1
        for _ in range(2):
            f = op_mul(f, lines[idx][0], lines[idx][1]); idx += 1
        assert idx == len(lines)
    gt = final_exponentiation(f)
    return gt == [1] + [0] * 11, gt
