#!/usr/bin/env python3
"""Generator (and CPU emulator) for the inline-PTX field kernels in fp_ptx.inc.

Why a generator: the 256-bit Montgomery multiplication that every kernel in this library bottoms out in is a ~300
instruction carry-chain program.  Writing it as ONE asm block per operation keeps the CC.CF carry flag inside a
single statement (no reliance on the compiler keeping separate asm statements in order), and lets this script
*execute the very same instruction list* on the CPU against Python big integers before any GPU time is spent
(`python gen_field_ptx.py --selftest`).

Scheme (8 x 32-bit limbs, R = 2^256): operand-scanning Montgomery with two interleaved accumulators.  "even" holds
64-bit columns at limb positions (0,1),(2,3),(4,5),(6,7), "odd" holds columns at (1,2),(3,4),(5,6),(7,8).  Every
32x32->64 partial product a[j]*b[i] is then a `mad.lo.cc / madc.hi.cc` pair on one aligned register pair, which
ptxas fuses into a single IMAD.WIDE.U32(.X) — half the issue slots of the textbook lo/hi formulation.  After each
b-limb the low limb is cancelled with q*m and the (even, odd) roles swap, which realises the divide-by-2^32.
Bounds: for inputs < 2m and m < 2^254 every intermediate total is < 2^288, so the odd chain never carries out and the
even chain's carry-out lands in odd[7]; the result before the final conditional subtraction is < 2m.
"""
import random
import sys

MASK = 0xFFFFFFFF

FQ = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
FR = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
# Pasta (the IPA decider's curve, halo2curves `pasta::pallas`): Pallas base field and scalar field.  Both are 255-bit;
# the bounds argument above is re-checked for them by the emulator (`--selftest` asserts that no carry is ever dropped).
PALLAS_P = 0x40000000000000000000000000000000224698FC094CF91B992D30ED00000001
PALLAS_Q = 0x40000000000000000000000000000000224698FC0994A8DD8C46EB2100000001


def limbs(x, n=8):
    return [(x >> (32 * i)) & MASK for i in range(n)]


class Prog:
    """A straight-line PTX program over named u32 registers with the CC.CF flag."""

    def __init__(self):
        self.ins = []      # (op, dst, [srcs])  srcs are register names or int immediates
        self.temps = []

    def tmp(self, name):
        if name not in self.temps:
            self.temps.append(name)
        return name

    def emit(self, op, dst, *srcs):
        self.ins.append((op, dst, list(srcs)))

    # ---- rendering -------------------------------------------------------------------------------------------
    def render(self, outs, ins_):
        """outs / ins_: ordered lists of external register names; returns the asm template body."""
        ext = {name: "%%%d" % i for i, name in enumerate(outs + ins_)}

        def r(x):
            if isinstance(x, int):
                return "0x%08x" % x
            return ext.get(x, x)
        lines = ["{"]
        decl = [t for t in self.temps if t not in ext]
        for k in range(0, len(decl), 12):
            lines.append(".reg .u32 " + ", ".join(decl[k:k + 12]) + ";")
        for op, dst, srcs in self.ins:
            lines.append("%s %s;" % (op, ", ".join([r(dst)] + [r(s) for s in srcs])))
        lines.append("}")
        return lines

    # ---- emulation -------------------------------------------------------------------------------------------
    def run(self, env):
        cf = 0
        env = dict(env)

        def v(x):
            return x if isinstance(x, int) else env[x]
        for op, dst, srcs in self.ins:
            s = [v(x) for x in srcs]
            base = op.split(".")[0]
            has_cc_out = ".cc" in op
            if base in ("add", "addc"):
                t = s[0] + s[1] + (cf if base == "addc" else 0)
            elif base in ("sub", "subc"):
                t = s[0] - s[1] - (cf if base == "subc" else 0)
                env[dst] = t & MASK
                if has_cc_out:
                    cf = 1 if t < 0 else 0      # CF = borrow
                continue
            elif base == "mul":
                p = s[0] * s[1]
                t = (p >> 32) if ".hi" in op else (p & MASK)
                env[dst] = t & MASK
                continue
            elif base in ("mad", "madc"):
                p = s[0] * s[1]
                part = (p >> 32) if ".hi" in op else (p & MASK)
                t = part + s[2] + (cf if base == "madc" else 0)
            elif base == "mov":
                env[dst] = s[0]
                continue
            elif base == "and":
                env[dst] = s[0] & s[1]
                continue
            elif base == "shl":
                env[dst] = (s[0] << s[1]) & MASK
                continue
            elif base == "shf":       # shf.l.wrap.b32 d, lo, hi, n : high word of (hi:lo) << n
                assert op == "shf.l.wrap.b32"
                env[dst] = ((((s[1] << 32) | s[0]) << (s[2] & 31)) >> 32) & MASK
                continue
            else:
                raise ValueError(op)
            env[dst] = t & MASK
            if has_cc_out:
                cf = t >> 32
                assert cf in (0, 1)
            elif t >> 32:
                # a dropped carry must never happen in these programs
                raise AssertionError("carry lost at %s %s" % (op, dst))
        return env


def gen_mont_mul(mod, a="a", b="b", out="r", reduce_final=True, b_const=None, second=None):
    """r = a * b / 2^256 mod m  (fully reduced if reduce_final, else < 2m).
    b_const: list of 8 limbs, each 0 or 1 — b is that CONSTANT and every product by it degenerates to a move / an add-with-carry
    (used for b = 1: r = a / 2^256 mod m, the conversion out of Montgomery form, at half the products of a multiplication).
    second = (c, d): r = (a * b + c * d) / 2^256 mod m — two products under ONE reduction (128 + 64 partial products instead of 256);
    inputs must be fully reduced (< m < 2^254) for the row sums to stay below 2^288 (asserted by the emulator)."""
    M = limbs(mod)
    m0inv = (-pow(mod, -1, 1 << 32)) & MASK
    p = Prog()
    A = [a + str(i) for i in range(8)]
    B = [b + str(i) for i in range(8)] if b_const is None else list(b_const)
    assert all(isinstance(x, str) or x in (0, 1) for x in B)
    E = [p.tmp("e%d" % i) for i in range(8)]
    O = [p.tmp("o%d" % i) for i in range(8)]
    mi = p.tmp("mi")

    def mul_n(acc, src_off, bi, src=A):
        for j in range(0, 8, 2):
            if isinstance(bi, int):          # constant 0 / 1
                if bi:
                    p.emit("mov.u32", acc[j], src[j + src_off])
                else:
                    p.emit("mov.u32", acc[j], 0)
                p.emit("mov.u32", acc[j + 1], 0)
                continue
            p.emit("mul.lo.u32", acc[j], src[j + src_off], bi)
            p.emit("mul.hi.u32", acc[j + 1], src[j + src_off], bi)

    def cmad_n(acc, srcs, bi, carry_out=True):
        # acc(64-bit columns) += srcs[0,2,4,6] * bi, one carry chain; CF left set on exit iff carry_out
        if isinstance(bi, int) and bi == 0:
            if carry_out:                     # nothing to add: the chain only has to leave CF = 0
                p.emit("add.cc.u32", acc[0], acc[0], 0)
            return
        assert not isinstance(bi, int)
        for k, j in enumerate(range(0, 8, 2)):
            p.emit("mad.lo.cc.u32" if k == 0 else "madc.lo.cc.u32", acc[j], bi, srcs[j], acc[j])
            last = (j == 6) and not carry_out
            p.emit("madc.hi.u32" if last else "madc.hi.cc.u32", acc[j + 1], bi, srcs[j], acc[j + 1])

    def madc_n_rshift(odd, bi):
        # odd <- (odd >> 64) + A[1,3,5,7] * bi + CF
        if isinstance(bi, int) and bi == 0:
            for j in range(0, 6):
                p.emit("addc.cc.u32", odd[j], odd[j + 2], 0)
            p.emit("addc.cc.u32", odd[6], 0, 0)
            p.emit("addc.u32", odd[7], 0, 0)
            return
        assert not isinstance(bi, int)
        for j in range(0, 6, 2):
            p.emit("madc.lo.cc.u32", odd[j], A[j + 1], bi, odd[j + 2])
            p.emit("madc.hi.cc.u32", odd[j + 1], A[j + 1], bi, odd[j + 3])
        p.emit("madc.lo.cc.u32", odd[6], A[7], bi, 0)
        p.emit("madc.hi.u32", odd[7], A[7], bi, 0)

    C2 = [second[0] + str(i) for i in range(8)] if second else None
    D2 = [second[1] + str(i) for i in range(8)] if second else None
    even, odd = E, O
    for i in range(8):
        bi = B[i]
        if i == 0:
            mul_n(odd, 1, bi)
            mul_n(even, 0, bi)
        else:
            p.emit("add.cc.u32", even[0], even[0], odd[1])
            madc_n_rshift(odd, bi)
            cmad_n(even, A, bi)
            p.emit("addc.u32", odd[7], odd[7], 0)
        if second:
            # + c * d_i into the same window: odd limbs of c on the odd columns (cannot carry out of the window), even limbs on the
            # even columns with the carry landing in odd[7]
            cmad_n(odd, C2[1:] + [0], D2[i], carry_out=False)
            cmad_n(even, C2, D2[i])
            p.emit("addc.u32", odd[7], odd[7], 0)
        p.emit("mul.lo.u32", mi, even[0], m0inv)
        # columns (1,2),(3,4),(5,6),(7,8): m1, m3, m5, m7; this chain cannot carry out (total < 2^288) — the
        # emulator asserts it
        cmad_n(odd, M[1:] + [0], mi, carry_out=False)
        cmad_n(even, M, mi)
        p.emit("addc.u32", odd[7], odd[7], 0)
        even, odd = odd, even
    # merge: result[k] = even[k] + odd[k+1]
    Rr = [p.tmp("t%d" % i) for i in range(8)]
    p.emit("add.cc.u32", Rr[0], even[0], odd[1])
    for k in range(1, 7):
        p.emit("addc.cc.u32", Rr[k], even[k], odd[k + 1])
    p.emit("addc.u32", Rr[7], even[7], 0)
    OUT = [out + str(i) for i in range(8)]
    if reduce_final:
        emit_cond_sub(p, Rr, M, OUT)
    else:
        for k in range(8):
            p.emit("mov.u32", OUT[k], Rr[k])
    if second:
        return p, A, B, C2, D2, OUT
    return p, A, B, OUT


def gen_mont_sqr(mod, a="a", out="r", reduce_final=True):
    """r = a * a / 2^256 mod m with 36 + 64 instead of 64 + 64 partial products (inputs < 2m < 2^255).
    Same interleaved operand scanning as gen_mont_mul, with every unordered pair of limbs assigned to the EARLIER row: row i multiplies
    a_i by C_i = a_i 2^(32 i) + 2 (a_{i+1} 2^(32 (i+1)) + ... + a_7 2^(32 7)), whose limbs are a_i, a_{i+1} << 1, then the limbs of 2a
    (d_j = a_j << 1 | a_{j-1} >> 31); limbs below i are zero and their products become carry propagation.  Row sums stay below 2^288
    (the emulator asserts that no carry is ever dropped)."""
    M = limbs(mod)
    m0inv = (-pow(mod, -1, 1 << 32)) & MASK
    p = Prog()
    A = [a + str(i) for i in range(8)]
    E = [p.tmp("e%d" % i) for i in range(8)]
    O = [p.tmp("o%d" % i) for i in range(8)]
    D = [None] + [p.tmp("d%d" % i) for i in range(1, 8)]     # limbs of 2a (d_8 = a_7 >> 31 = 0 for a < 2^255)
    S = [None] + [p.tmp("h%d" % i) for i in range(1, 8)]     # a_j << 1 (the limb right above a_i in C_i)
    mi = p.tmp("mi")
    for j in range(1, 8):
        p.emit("shf.l.wrap.b32", D[j], A[j - 1], A[j], 1)
        p.emit("shl.b32", S[j], A[j], 1)

    def row_vec(i):
        v = [0] * 8
        v[i] = A[i]
        if i + 1 < 8:
            v[i + 1] = S[i + 1]
        for j in range(i + 2, 8):
            v[j] = D[j]
        return v

    def cmad_sparse(acc, srcs, bi, carry_out=True):
        # acc(64-bit columns at limbs (0,1),(2,3),(4,5),(6,7)) += srcs[0,2,4,6] * bi; zero entries only pass the carry on
        started = False
        for j in range(0, 8, 2):
            last = (j == 6) and not carry_out
            if isinstance(srcs[j], int):
                assert srcs[j] == 0
                if started:
                    p.emit("addc.cc.u32", acc[j], acc[j], 0)
                    p.emit("addc.u32" if last else "addc.cc.u32", acc[j + 1], acc[j + 1], 0)
                continue
            p.emit("madc.lo.cc.u32" if started else "mad.lo.cc.u32", acc[j], bi, srcs[j], acc[j])
            p.emit("madc.hi.u32" if last else "madc.hi.cc.u32", acc[j + 1], bi, srcs[j], acc[j + 1])
            started = True
        if not started and carry_out:
            p.emit("add.cc.u32", acc[0], acc[0], 0)          # leave CF = 0 for the addc that follows

    def cmad_const(acc, consts, bi, carry_out=True):
        for k, j in enumerate(range(0, 8, 2)):
            p.emit("mad.lo.cc.u32" if k == 0 else "madc.lo.cc.u32", acc[j], bi, consts[j], acc[j])
            last = (j == 6) and not carry_out
            p.emit("madc.hi.u32" if last else "madc.hi.cc.u32", acc[j + 1], bi, consts[j], acc[j + 1])

    def rshift_mad_sparse(odd, srcs, bi):
        # odd <- (odd >> 64) + srcs[1,3,5,7] * bi + CF
        for j in range(0, 6, 2):
            if isinstance(srcs[j + 1], int):
                p.emit("addc.cc.u32", odd[j], odd[j + 2], 0)
                p.emit("addc.cc.u32", odd[j + 1], odd[j + 3], 0)
            else:
                p.emit("madc.lo.cc.u32", odd[j], srcs[j + 1], bi, odd[j + 2])
                p.emit("madc.hi.cc.u32", odd[j + 1], srcs[j + 1], bi, odd[j + 3])
        if isinstance(srcs[7], int):
            p.emit("addc.cc.u32", odd[6], 0, 0)
            p.emit("addc.u32", odd[7], 0, 0)
        else:
            p.emit("madc.lo.cc.u32", odd[6], srcs[7], bi, 0)
            p.emit("madc.hi.u32", odd[7], srcs[7], bi, 0)

    even, odd = E, O
    for i in range(8):
        bi = A[i]
        v = row_vec(i)
        if i == 0:
            for j in range(0, 8, 2):
                p.emit("mul.lo.u32", odd[j], v[j + 1], bi)
                p.emit("mul.hi.u32", odd[j + 1], v[j + 1], bi)
            for j in range(0, 8, 2):
                p.emit("mul.lo.u32", even[j], v[j], bi)
                p.emit("mul.hi.u32", even[j + 1], v[j], bi)
        else:
            p.emit("add.cc.u32", even[0], even[0], odd[1])
            rshift_mad_sparse(odd, v, bi)
            cmad_sparse(even, v, bi)
            p.emit("addc.u32", odd[7], odd[7], 0)
        p.emit("mul.lo.u32", mi, even[0], m0inv)
        cmad_const(odd, M[1:] + [0], mi, carry_out=False)
        cmad_const(even, M, mi)
        p.emit("addc.u32", odd[7], odd[7], 0)
        even, odd = odd, even
    Rr = [p.tmp("t%d" % i) for i in range(8)]
    p.emit("add.cc.u32", Rr[0], even[0], odd[1])
    for k in range(1, 7):
        p.emit("addc.cc.u32", Rr[k], even[k], odd[k + 1])
    p.emit("addc.u32", Rr[7], even[7], 0)
    OUT = [out + str(i) for i in range(8)]
    if reduce_final:
        emit_cond_sub(p, Rr, M, OUT)
    else:
        for k in range(8):
            p.emit("mov.u32", OUT[k], Rr[k])
    return p, A, OUT


def emit_cond_sub(p, X, M, OUT):
    """OUT = X - M if X >= M else X   (branch-free: add back (M & borrow_mask))."""
    S = [p.tmp("s%d" % i) for i in range(8)]
    bm = p.tmp("bm")
    p.emit("sub.cc.u32", S[0], X[0], M[0])
    for k in range(1, 8):
        p.emit("subc.cc.u32", S[k], X[k], M[k])
    p.emit("subc.u32", bm, 0, 0)          # 0 or 0xffffffff (borrow)
    # OUT = S + (M & bm)
    for k in range(8):
        mk = p.tmp("mk%d" % k)
        p.emit("and.b32", mk, bm, M[k])
    p.emit("add.cc.u32", OUT[0], S[0], "mk0")
    for k in range(1, 7):
        p.emit("addc.cc.u32", OUT[k], S[k], "mk%d" % k)
    p.emit("addc.u32.nocheck", OUT[7], S[7], "mk7")


def gen_add_mod(mod):
    """r = a + b mod m, inputs < m."""
    M = limbs(mod)
    p = Prog()
    A = ["a%d" % i for i in range(8)]
    B = ["b%d" % i for i in range(8)]
    T = [p.tmp("t%d" % i) for i in range(8)]
    p.emit("add.cc.u32", T[0], A[0], B[0])
    for k in range(1, 7):
        p.emit("addc.cc.u32", T[k], A[k], B[k])
    p.emit("addc.u32", T[7], A[7], B[7])
    OUT = ["r%d" % i for i in range(8)]
    emit_cond_sub(p, T, M, OUT)
    return p, A, B, OUT


def gen_sub_mod(mod):
    """r = a - b mod m, inputs < m."""
    M = limbs(mod)
    p = Prog()
    A = ["a%d" % i for i in range(8)]
    B = ["b%d" % i for i in range(8)]
    S = [p.tmp("s%d" % i) for i in range(8)]
    bm = p.tmp("bm")
    p.emit("sub.cc.u32", S[0], A[0], B[0])
    for k in range(1, 8):
        p.emit("subc.cc.u32", S[k], A[k], B[k])
    p.emit("subc.u32", bm, 0, 0)
    OUT = ["r%d" % i for i in range(8)]
    for k in range(8):
        p.emit("and.b32", p.tmp("mk%d" % k), bm, M[k])
    p.emit("add.cc.u32", OUT[0], S[0], "mk0")
    for k in range(1, 7):
        p.emit("addc.cc.u32", OUT[k], S[k], "mk%d" % k)
    p.emit("addc.u32.nocheck", OUT[7], S[7], "mk7")
    return p, A, B, OUT


# `.nocheck` is an emulator-only suffix: the wrap-around of that last addc is intended (two's complement add-back).
def _strip(p):
    q = Prog()
    q.temps = p.temps
    q.ins = [(op.replace(".nocheck", ""), d, s) for op, d, s in p.ins]
    return q


def _emul(p, env):
    q = Prog()
    q.temps = p.temps
    fixed = []
    for op, d, s in p.ins:
        if op.endswith(".nocheck"):
            fixed.append((op.replace(".nocheck", "").replace("addc.u32", "addc.cc.u32"), d, s))
        else:
            fixed.append((op, d, s))
    q.ins = fixed
    return q.run(env)


def env_of(prefix, x):
    return {prefix + str(i): l for i, l in enumerate(limbs(x))}


def val_of(env, names):
    return sum(env[n] << (32 * i) for i, n in enumerate(names))


def selftest(iters=2000):
    rnd = random.Random(7)
    Rm = 1 << 256
    for name, mod in (("fq", FQ), ("fr", FR), ("pallas_p", PALLAS_P), ("pallas_q", PALLAS_Q)):
        Rinv = pow(Rm, -1, mod)
        pm, A, B, OUT = gen_mont_mul(mod)
        pl, _, _, OUTL = gen_mont_mul(mod, reduce_final=False)
        pa, _, _, OA = gen_add_mod(mod)
        ps, _, _, OS = gen_sub_mod(mod)
        pf, _, _, OF = gen_mont_mul(mod, b_const=[1, 0, 0, 0, 0, 0, 0, 0])
        pq, _, OQ = gen_mont_sqr(mod)
        p2, _, _, _, _, O2 = gen_mont_mul(mod, second=("c", "d"))
        edge = [0, 1, mod - 1, mod - 2, (1 << 254) % mod, Rm % mod]
        cases = [(x, y) for x in edge for y in edge] + [(rnd.randrange(mod), rnd.randrange(mod)) for _ in range(iters)]
        for x, y in cases:
            env = {}
            env.update(env_of("a", x)); env.update(env_of("b", y))
            assert val_of(_emul(pm, env), OUT) == x * y * Rinv % mod, (name, "mul", hex(x), hex(y))
            assert val_of(_emul(pa, env), OA) == (x + y) % mod, (name, "add")
            assert val_of(_emul(ps, env), OS) == (x - y) % mod, (name, "sub")
            assert val_of(_emul(pf, env), OF) == x * Rinv % mod, (name, "from_mont", hex(x))
            assert val_of(_emul(pq, env), OQ) == x * x * Rinv % mod, (name, "sqr", hex(x))
            if mod < (1 << 254):      # two products under one reduction (BN254 only: the row sums need m < 2^254)
                for u, w in ((y, x), (mod - 1, mod - 1), (rnd.randrange(mod), rnd.randrange(mod))):
                    env2 = dict(env); env2.update(env_of("c", u)); env2.update(env_of("d", w))
                    assert val_of(_emul(p2, env2), O2) == (x * y + u * w) * Rinv % mod, (name, "mul2add")
        # lazy variant: inputs anywhere below 2m, output < 2m and congruent (BN254 only: needs m < 2^254)
        pql, _, OQL = gen_mont_sqr(mod, reduce_final=False)
        for _ in range(iters if mod < (1 << 254) else 0):
            x = rnd.randrange(2 * mod)
            got = val_of(_emul(pql, env_of("a", x)), OQL)
            assert got < 2 * mod and got % mod == x * x * Rinv % mod, (name, "lazy sqr")
        for _ in range(iters if mod < (1 << 254) else 0):
            x, y = rnd.randrange(2 * mod), rnd.randrange(2 * mod)
            env = {}
            env.update(env_of("a", x)); env.update(env_of("b", y))
            got = val_of(_emul(pl, env), OUTL)
            assert got < 2 * mod and got % mod == x * y * Rinv % mod, (name, "lazy mul")
        print("selftest %s: %d mul/sqr/add/sub/from_mont cases ok (%d PTX instructions per mul, %d per sqr, %d per from_mont)"
              % (name, len(cases), len(pm.ins), len(pq.ins), len(pf.ins)))


def c_macro(name, lines):
    body = " \\\n".join('    "%s\\n\\t"' % ln for ln in lines)
    return "#define %s \\\n%s\n" % (name, body)


def main():
    if "--selftest" in sys.argv:
        selftest()
        return
    pallas = "--curve=pallas" in sys.argv
    out = ["// GENERATED by gen_field_ptx.py — do not edit.  Regenerate: python gen_field_ptx.py %s> %s" % (("--curve=pallas ", "fp_ptx_pallas.inc") if pallas else ("", "fp_ptx.inc")),
           "// Operand order of every block: %0..%7 = r[0..7] (out), %8..%15 = a[0..7], %16..%23 = b[0..7].",
           "// Verified on the CPU by `python gen_field_ptx.py --selftest` (PTX emulator vs Python big-ints).", ""]
    if pallas:
        out.insert(1, "// Pallas build: FQ = the Pallas BASE field, FR = the Pallas SCALAR field (same macro names, so every kernel is curve-agnostic).")
    for name, mod in ((("FQ", PALLAS_P), ("FR", PALLAS_Q)) if pallas else (("FQ", FQ), ("FR", FR))):
        R = (1 << 256) % mod
        for tag, val in (("MOD", mod), ("ONE", R), ("R2", R * R % mod), ("R3", R * R * R % mod)):
            out.append("#define SNARKV_%s_%s_LIMBS {%s}" % (name, tag, ", ".join("0x%08xu" % l for l in limbs(val))))
        out.append("#define SNARKV_%s_M0INV 0x%08xu" % (name, (-pow(mod, -1, 1 << 32)) & MASK))
        out.append("")
        ops = [("MUL", gen_mont_mul(mod)), ("ADD", gen_add_mod(mod)), ("SUB", gen_sub_mod(mod))]
        if not pallas:
            ops.insert(1, ("MUL_LAZY", gen_mont_mul(mod, reduce_final=False)))
        for tag, (p, A, B, OUT) in ops:
            out.append(c_macro("SNARKV_PTX_%s_%s" % (name, tag), _strip(p).render(OUT, A + B)))
        # r = a / 2^256 mod m (out of Montgomery form): operands %0..%7 = r, %8..%15 = a
        p, A, _, OUT = gen_mont_mul(mod, b_const=[1, 0, 0, 0, 0, 0, 0, 0])
        out.append(c_macro("SNARKV_PTX_%s_FROM_MONT" % name, _strip(p).render(OUT, A)))
        if not pallas:
            # r = (a * b + c * d) / 2^256 mod m, one reduction: operands %0..%7 = r, %8..%15 = a, %16..%23 = b, %24..%31 = c, %32..%39 = d
            p, A, B, C2, D2, OUT = gen_mont_mul(mod, second=("c", "d"))
            out.append(c_macro("SNARKV_PTX_%s_MUL2ADD" % name, _strip(p).render(OUT, A + B + C2 + D2)))
        # r = a * a / 2^256 mod m with 36 + 64 partial products: operands %0..%7 = r, %8..%15 = a
        p, A, OUT = gen_mont_sqr(mod)
        out.append(c_macro("SNARKV_PTX_%s_SQR" % name, _strip(p).render(OUT, A)))
    sys.stdout.write("\n".join(out))


if __name__ == "__main__":
    main()
