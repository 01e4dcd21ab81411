"""Host mirror of the per-proof PLONK scalar evaluation of snark-verifier and its compiler to the
straight-line Fr register program that `snarkv_fr_program_eval_batch` (csrc/fr_program.cu) runs for a whole batch of proofs.

Mirrored reference items (paths relative to snark-verifier/src):
  util/arithmetic.rs:83-160        root_of_unity, Rotation, Domain::{new, rotate_scalar}
  verifier/plonk/protocol.rs:186-197  CommonPolynomial::{Identity, Lagrange(i)}
  verifier/plonk/protocol.rs:211-283  CommonPolynomialEvaluation::{new, evaluate}: z^n, z^n - 1, 1/(z^n - 1), L_i(z)
  verifier/plonk/protocol.rs:304-392  Query, Expression and Expression::evaluate (the eight-closure fold, DistributePowers)
  verifier/plonk/proof.rs:298-303     quotient evaluation = numerator * zn_minus_one_inv (no linearization)
  verifier/plonk/proof.rs:306-349     PlonkProof::evaluations: instance evaluations sum_j instance[j] * L_{j - rotation}(z)
  loader.rs:52-69                     LoadedScalar::pow_const (exact square-and-multiply order)
  loader.rs:255-262                   ScalarLoader::batch_invert (zero stays zero)

The protocol of a batch is fixed, so the expression tree is the same for every proof: `Expression.evaluate` is run ONCE on the
host with closures that emit instructions instead of computing (exactly how the reference runs it over EVM / Halo2 loaders), and
the resulting program is executed by one GPU thread per proof.  The linearized (Msm-valued) numerator of
`LinearizationStrategy::*` stays on the host: this module covers the scalar path (`linearization: None`, what the halo2 system
produces).
"""
from dataclasses import dataclass
from typing import Dict, List, Sequence, Tuple

R_MODULUS = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
FR_S = 28                                                                          # Fr::S
FR_ROOT_OF_UNITY = 0x03DDB9F5166D18B798865EA93DD31F743215CF6DD39329C8D34F1ED960C37C9C  # Fr::ROOT_OF_UNITY = 7^((r-1)/2^28)

OP_INPUT, OP_CONST, OP_ADD, OP_SUB, OP_MUL, OP_NEG, OP_INV, OP_NZ, OP_KEEPZ = range(9)   # include/snarkv_cuda.h SNARKV_FR_OP_*


def root_of_unity(k: int) -> int:
    """util/arithmetic.rs:83-90"""
    assert k <= FR_S
    w = FR_ROOT_OF_UNITY
    for _ in range(FR_S - k):
        w = w * w % R_MODULUS
    return w


@dataclass(frozen=True, order=True)
class Rotation:
    """util/arithmetic.rs:95-120"""
    value: int = 0


class Domain:
    """util/arithmetic.rs:123-160"""

    def __init__(self, k: int, gen: int = None):
        self.k = k
        self.n = 1 << k
        self.gen = root_of_unity(k) if gen is None else gen
        self.n_inv = pow(self.n, -1, R_MODULUS)
        self.gen_inv = pow(self.gen, -1, R_MODULUS)

    def rotate_scalar(self, scalar: int, rotation: Rotation) -> int:
        r = rotation.value
        if r == 0:
            return scalar
        if r > 0:
            return scalar * pow(self.gen, r, R_MODULUS) % R_MODULUS
        return scalar * pow(self.gen_inv, -r, R_MODULUS) % R_MODULUS


@dataclass(frozen=True, order=True)
class Query:
    """protocol.rs:304-320"""
    poly: int
    rotation: Rotation = Rotation(0)


@dataclass(frozen=True)
class CommonPolynomial:
    """protocol.rs:186-197: Identity, or Lagrange(i)"""
    kind: str            # "identity" | "lagrange"
    index: int = 0

    @staticmethod
    def identity():
        return CommonPolynomial("identity")

    @staticmethod
    def lagrange(i: int):
        return CommonPolynomial("lagrange", i)


class Expression:
    """protocol.rs:322-334.  Variants are built with the static constructors; `evaluate` is the reference's generic fold."""

    __slots__ = ("tag", "args")

    def __init__(self, tag, *args):
        self.tag, self.args = tag, args

    # -- constructors --------------------------------------------------------------------------------------------------
    @staticmethod
    def constant(c: int): return Expression("constant", c % R_MODULUS)
    @staticmethod
    def common_polynomial(p: CommonPolynomial): return Expression("common", p)
    @staticmethod
    def polynomial(q: Query): return Expression("poly", q)
    @staticmethod
    def challenge(i: int): return Expression("challenge", i)
    @staticmethod
    def negated(a): return Expression("neg", a)
    @staticmethod
    def sum(a, b): return Expression("sum", a, b)
    @staticmethod
    def product(a, b): return Expression("product", a, b)
    @staticmethod
    def scaled(a, c: int): return Expression("scaled", a, c % R_MODULUS)
    @staticmethod
    def distribute_powers(exprs: Sequence["Expression"], scalar: "Expression"): return Expression("powers", tuple(exprs), scalar)

    def __add__(self, o): return Expression.sum(self, o)
    def __mul__(self, o): return Expression.product(self, o) if isinstance(o, Expression) else Expression.scaled(self, o)
    def __neg__(self): return Expression.negated(self)
    def __sub__(self, o): return Expression.sum(self, Expression.negated(o))

    # -- protocol.rs:336-392 -----------------------------------------------------------------------------------------------
    def evaluate(self, constant, common_poly, poly, challenge, negated, sum, product, scaled):
        ev = lambda e: e.evaluate(constant, common_poly, poly, challenge, negated, sum, product, scaled)
        t, a = self.tag, self.args
        if t == "constant":
            return constant(a[0])
        if t == "common":
            return common_poly(a[0])
        if t == "poly":
            return poly(a[0])
        if t == "challenge":
            return challenge(a[0])
        if t == "neg":
            return negated(ev(a[0]))
        if t == "sum":
            x = ev(a[0]); y = ev(a[1])
            return sum(x, y)
        if t == "product":
            x = ev(a[0]); y = ev(a[1])
            return product(x, y)
        if t == "scaled":
            return scaled(ev(a[0]), a[1])
        exprs, scalar = a                                  # DistributePowers: Horner in `scalar`, first expression = highest power
        assert len(exprs) > 0
        if len(exprs) == 1:
            return ev(exprs[0])
        acc = ev(exprs[0])
        s = ev(scalar)
        for e in exprs[1:]:
            acc = sum(product(acc, s), ev(e))
        return acc

    def used_langrange(self) -> set:
        """protocol.rs:417-435"""
        return self.evaluate(lambda c: set(), lambda p: {p.index} if p.kind == "lagrange" else set(), lambda q: set(), lambda i: set(),
                             lambda a: a, lambda a, b: a | b, lambda a, b: a | b, lambda a, c: a)

    def used_query(self) -> set:
        """protocol.rs:437-455"""
        return self.evaluate(lambda c: set(), lambda p: set(), lambda q: {q}, lambda i: set(),
                             lambda a: a, lambda a, b: a | b, lambda a, b: a | b, lambda a, c: a)

    def to_tuple(self):
        """Plain nested tuples (what the test oracle consumes; it shares no code with this module)."""
        t, a = self.tag, self.args
        if t == "common":
            return ("common", a[0].kind, a[0].index)
        if t == "poly":
            return ("poly", a[0].poly, a[0].rotation.value)
        if t in ("constant", "challenge"):
            return (t, a[0])
        if t == "neg":
            return ("neg", a[0].to_tuple())
        if t in ("sum", "product"):
            return (t, a[0].to_tuple(), a[1].to_tuple())
        if t == "scaled":
            return ("scaled", a[0].to_tuple(), a[1])
        return ("powers", tuple(e.to_tuple() for e in a[0]), a[1].to_tuple())


# ----------------------------------------------------------------------------------------------------------------------
# program builder: the "loader" whose scalars are virtual registers
# ----------------------------------------------------------------------------------------------------------------------
class Program:
    """A straight-line register program: `instrs` (op, dst, a, b) over `n_regs` registers, `consts` (ints mod r), `n_inputs`
    per-proof input slots and the `outputs` registers, as snarkv_fr_program_eval_batch takes them."""

    def __init__(self, instrs, n_regs, consts, n_inputs, outputs, output_names):
        self.instrs, self.n_regs, self.consts, self.n_inputs = instrs, n_regs, consts, n_inputs
        self.outputs, self.output_names = outputs, output_names

    def op_histogram(self) -> Dict[str, int]:
        names = ["input", "const", "add", "sub", "mul", "neg", "inv", "nz", "keepz"]
        h = {k: 0 for k in names}
        for op, _, _, _ in self.instrs:
            h[names[op]] += 1
        return h


class ProgramBuilder:
    """Emits SSA values; `finish` maps them onto a small register file by liveness (a value's register is reused after its last
    use), which keeps the device register file — n_regs x m x 32 B — L2-sized."""

    def __init__(self):
        self.ssa: List[Tuple[int, int, int]] = []     # (op, a, b): a, b are SSA ids, or the slot for INPUT / CONST
        self.consts: List[int] = []
        self._const_ix: Dict[int, int] = {}
        self._const_val: Dict[int, int] = {}
        self._input_val: Dict[int, int] = {}
        self.n_inputs = 0

    def _emit(self, op, a=0, b=0) -> int:
        self.ssa.append((op, a, b))
        return len(self.ssa) - 1

    def input(self, slot: int) -> int:
        """Per-proof value `slot` of the input row (loaded once)."""
        if slot not in self._input_val:
            self._input_val[slot] = self._emit(OP_INPUT, slot)
            self.n_inputs = max(self.n_inputs, slot + 1)
        return self._input_val[slot]

    def const(self, c: int) -> int:
        """ScalarLoader::load_const (loaded once per distinct value)."""
        c %= R_MODULUS
        if c not in self._const_ix:
            self._const_ix[c] = len(self.consts)
            self.consts.append(c)
        ix = self._const_ix[c]
        if ix not in self._const_val:
            self._const_val[ix] = self._emit(OP_CONST, ix)
        return self._const_val[ix]

    def add(self, a, b): return self._emit(OP_ADD, a, b)
    def sub(self, a, b): return self._emit(OP_SUB, a, b)
    def mul(self, a, b): return self._emit(OP_MUL, a, b)
    def neg(self, a): return self._emit(OP_NEG, a)
    def inv(self, a): return self._emit(OP_INV, a)
    def nz(self, a): return self._emit(OP_NZ, a)
    def keepz(self, a, b): return self._emit(OP_KEEPZ, a, b)

    def batch_invert(self, values: Sequence[int]) -> List[int]:
        """util/arithmetic.rs:47-74 (`batch_invert`, what NativeLoader's `ScalarLoader::batch_invert` calls) as straight-line code:
        running products of the non-zero values, ONE inversion, back-substitution; a zero value stays zero.  The data-dependent
        `filter(!is_zero)` becomes two selects: NZ feeds a 1 into the product in place of a zero, KEEPZ restores the zero."""
        if not values:
            return []
        nzv = [self.nz(v) for v in values]
        products = [nzv[0]]
        for v in nzv[1:]:
            products.append(self.mul(products[-1], v))
        all_inv = self.inv(products[-1])
        out = [None] * len(values)
        for i in range(len(values) - 1, -1, -1):
            inv_i = self.mul(all_inv, products[i - 1]) if i > 0 else all_inv
            if i > 0:
                all_inv = self.mul(all_inv, nzv[i])
            out[i] = self.keepz(inv_i, values[i])
        return out

    def pow_const(self, a: int, exp: int) -> int:
        """loader.rs:52-69, the same square-and-multiply order."""
        assert exp > 0
        base = a
        while exp & 1 == 0:
            base = self.mul(base, base)
            exp >>= 1
        acc = base
        while exp > 1:
            exp >>= 1
            base = self.mul(base, base)
            if exp & 1:
                acc = self.mul(acc, base)
        return acc

    def finish(self, outputs: Sequence[int], names: Sequence[str] = None) -> Program:
        n = len(self.ssa)
        last = list(range(n))                        # last SSA id that reads each value
        for i, (op, a, b) in enumerate(self.ssa):
            if op in (OP_ADD, OP_SUB, OP_MUL, OP_KEEPZ):
                last[a] = max(last[a], i); last[b] = max(last[b], i)
            elif op in (OP_NEG, OP_INV, OP_NZ):
                last[a] = max(last[a], i)
        for o in outputs:
            last[o] = n                              # outputs live to the end
        free: List[int] = []
        reg = [0] * n
        n_regs = 0
        expire: Dict[int, List[int]] = {}
        instrs = []
        for i, (op, a, b) in enumerate(self.ssa):
            if op in (OP_INPUT, OP_CONST):
                ra, rb = a, 0
            elif op in (OP_NEG, OP_INV, OP_NZ):
                ra, rb = reg[a], 0
            else:
                ra, rb = reg[a], reg[b]
            # operands whose last use is this instruction release their registers first: dst may reuse them (the kernel reads
            # both operands before it writes)
            for v in expire.pop(i, []):
                free.append(reg[v])
            if free:
                r = free.pop()
            else:
                r = n_regs
                n_regs += 1
            reg[i] = r
            if last[i] == i:                         # never read (dead value): release at once
                free.append(r)
            elif last[i] < n:
                expire.setdefault(last[i], []).append(i)
            instrs.append((op, r, ra, rb))
        return Program(instrs, max(n_regs, 1), list(self.consts), self.n_inputs, [reg[o] for o in outputs],
                       list(names) if names else ["out%d" % k for k in range(len(outputs))])


# ----------------------------------------------------------------------------------------------------------------------
# protocol.rs:211-283 and proof.rs:298-349 over the builder
# ----------------------------------------------------------------------------------------------------------------------
class CommonPolynomialEvaluation:
    """protocol.rs:199-283 with `ProgramBuilder` values: `new` collects the fractions, then — as PlonkVerifier does through
    `denoms()` + `L::batch_invert` + `evaluate()` (verifier/plonk.rs, protocol.rs:263-282) — all denominators of a proof are
    inverted together (Lagrange denominators in index order, then z^n - 1) and the fractions are evaluated."""

    def __init__(self, b: ProgramBuilder, domain: Domain, langranges, z: int):
        self.b = b
        self.zn = b.pow_const(z, domain.n)
        lang = sorted(set(langranges))
        one = b.const(1)
        self.zn_minus_one = b.sub(self.zn, one)
        n_inv = b.const(domain.n_inv)
        numer = b.mul(self.zn_minus_one, n_inv)
        self.identity = z
        numers, denoms = [], []
        for i in lang:
            omega = b.const(domain.rotate_scalar(1, Rotation(i)))
            numers.append(b.mul(numer, omega))                      # Fraction::new(numer * omega, z - omega)
            denoms.append(b.sub(z, omega))
        inv = b.batch_invert(denoms + [self.zn_minus_one])          # denoms(): lagrange values, then zn_minus_one_inv
        self.zn_minus_one_inv = inv[-1]                             # Fraction::one_over(zn_minus_one), evaluated
        self.lagrange = {i: b.mul(nu, dinv) for i, nu, dinv in zip(lang, numers, inv)}   # Fraction::evaluate: numer * denom^-1

    def get(self, poly: CommonPolynomial) -> int:
        return self.identity if poly.kind == "identity" else self.lagrange[poly.index]


@dataclass
class QuotientProtocol:
    """The slice of `PlonkProtocol` (protocol.rs:22-67) the scalar path reads."""
    domain: Domain
    num_preprocessed: int
    num_instance: List[int]              # protocol.num_instance
    evaluations: List[Query]             # protocol.evaluations: the queries whose evaluations the proof carries, in order
    num_challenge: int                   # sum(protocol.num_challenge)
    numerator: Expression                # protocol.quotient.numerator

    def input_layout(self) -> Dict[str, int]:
        """Per-proof input row: [z | challenges | evaluations (protocol.evaluations order) | instances (column-major)]."""
        off = {"z": 0, "challenges": 1}
        off["evaluations"] = 1 + self.num_challenge
        off["instances"] = off["evaluations"] + len(self.evaluations)
        off["total"] = off["instances"] + sum(self.num_instance)
        return off


def compile_quotient_evaluation(p: QuotientProtocol) -> Program:
    """The scalar half of PlonkProof::{evaluations, commitments} for `linearization: None` and no instance committing key:
    outputs [quotient evaluation, z^n, z^n - 1, 1/(z^n - 1), instance evaluations...] per proof."""
    b = ProgramBuilder()
    lay = p.input_layout()
    z = b.input(lay["z"])
    offset = p.num_preprocessed
    inst_range = range(offset, offset + len(p.num_instance))
    inst_queries = sorted(q for q in p.numerator.used_query() if q.poly in inst_range)
    # PlonkProtocol::langranges (protocol.rs:77-106): the numerator's own Lagrange indices plus ONE range for the instance
    # evaluations, -max_rotation .. max_instance_len + |min_rotation| (min / max folded from (0, 0) over the instance queries)
    lang = set(p.numerator.used_langrange())
    max_inst = max(p.num_instance) if p.num_instance else 0
    min_rot = max_rot = 0
    for q in inst_queries:
        if q.rotation.value < min_rot:
            min_rot = q.rotation.value
        elif q.rotation.value > max_rot:
            max_rot = q.rotation.value
    lang |= set(range(-max_rot, max_inst + abs(min_rot)))
    cpe = CommonPolynomialEvaluation(b, p.domain, lang, z)
    evals: Dict[Query, int] = {}
    inst_base = [lay["instances"] + sum(p.num_instance[:k]) for k in range(len(p.num_instance))]
    inst_eval_regs = []
    for q in inst_queries:                                           # proof.rs:313-334 (loader.sum_products)
        col = q.poly - offset
        acc = None
        for j in range(p.num_instance[col]):
            term = b.mul(b.input(inst_base[col] + j), cpe.get(CommonPolynomial.lagrange(j - q.rotation.value)))
            acc = term if acc is None else b.add(acc, term)
        if acc is None:
            acc = b.const(0)
        evals[q] = acc
        inst_eval_regs.append(acc)
    for k, q in enumerate(p.evaluations):                           # proof.rs:336-346
        evals[q] = b.input(lay["evaluations"] + k)

    def poly(q):
        if q not in evals:
            raise KeyError("Missing query %r" % (q,))                # Error::InvalidProtocol("Missing query ..")
        return evals[q]

    def challenge(i):
        if i >= p.num_challenge:
            raise KeyError("Missing challenge %d" % i)
        return b.input(lay["challenges"] + i)
    numerator = p.numerator.evaluate(lambda c: b.const(c), cpe.get, poly, challenge, b.neg, b.add, b.mul,
                                     lambda a, c: b.mul(a, b.const(c)))
    quotient_eval = b.mul(numerator, cpe.zn_minus_one_inv)           # proof.rs:298-303
    outs = [quotient_eval, cpe.zn, cpe.zn_minus_one, cpe.zn_minus_one_inv] + inst_eval_regs
    names = ["quotient_eval", "zn", "zn_minus_one", "zn_minus_one_inv"] + ["instance_eval_%d" % k for k in range(len(inst_eval_regs))]
    b.n_inputs = lay["total"]            # the row stride is the layout's, even when the numerator leaves some slots unread
    return b.finish(outs, names)


def pack_program(prog: Program):
    """ctypes-ready arrays: (instr uint32[n][4], consts bytes LE canonical, out_regs uint32[n_out])."""
    import numpy as np
    ins = np.asarray(prog.instrs, dtype=np.uint32).reshape(-1, 4)
    consts = b"".join(c.to_bytes(32, "little") for c in prog.consts)
    outs = np.asarray(prog.outputs, dtype=np.uint32)
    return ins, consts, outs


# ----------------------------------------------------------------------------------------------------------------------
# synthetic protocol of StandardPlonk shape (tests / bench): what system/halo2.rs compiles for the reference's StandardPlonk
# test circuit, built by hand — gate + permutation argument over (a, b, c) with zero-knowledge rows, no lookups
# ----------------------------------------------------------------------------------------------------------------------
FR_DELTA = pow(7, 1 << FR_S, R_MODULUS)   # Fr::DELTA = MULTIPLICATIVE_GENERATOR^(2^S)


def standard_plonk_like_protocol(k: int, num_instance: int = 1, blinding_factors: int = 5) -> QuotientProtocol:
    """Polynomial indices: 0-4 fixed (q_a, q_b, q_c, q_ab, constant), 5-7 permutation sigmas, 8 instance, 9-11 advice (a, b, c),
    12 permutation z.  Challenges: 0 theta (unused), 1 beta, 2 gamma, 3 alpha.  Constraints as system/halo2.rs:520-660 lays them
    out: gate; l_0 (1 - z); l_last (z^2 - z); l_active (z(wX) prod(p_i + beta sigma_i + gamma) - z prod(p_i + beta delta^i X + gamma));
    numerator = DistributePowers(constraints, alpha)."""
    E, Q, CP = Expression, Query, CommonPolynomial
    poly = lambda i, r=0: E.polynomial(Q(i, Rotation(r)))
    q_a, q_b, q_c, q_ab, constant = (poly(i) for i in range(5))
    sigmas = [poly(5 + i) for i in range(3)]
    instance = poly(8)
    advice = [poly(9 + i) for i in range(3)]
    a, b, c = advice
    z, z_omega = poly(12), poly(12, 1)
    beta, gamma, alpha = E.challenge(1), E.challenge(2), E.challenge(3)
    one = E.constant(1)
    rotation_last = -(blinding_factors + 1)
    l_0 = E.common_polynomial(CP.lagrange(0))
    l_last = E.common_polynomial(CP.lagrange(rotation_last))
    l_blind = None
    for i in range(rotation_last + 1, 0):
        t = E.common_polynomial(CP.lagrange(i))
        l_blind = t if l_blind is None else l_blind + t
    l_active = one - (l_last + l_blind)
    identity = E.common_polynomial(CP.identity())
    gate = q_a * a + q_b * b + q_c * c + q_ab * a * b + constant + instance
    left = z_omega
    for p, s in zip(advice, sigmas):
        left = left * (p + beta * s + gamma)
    right = z
    delta = 1
    for p in advice:
        right = right * (p + beta * E.constant(delta) * identity + gamma)
        delta = delta * FR_DELTA % R_MODULUS
    constraints = [gate, l_0 * (one - z), l_last * (z * z - z), l_active * (left - right)]
    evaluations = [Q(i, Rotation(0)) for i in range(8)] + [Q(9 + i, Rotation(0)) for i in range(3)] + [Q(12, Rotation(0)), Q(12, Rotation(1))]
    return QuotientProtocol(domain=Domain(k), num_preprocessed=8, num_instance=[num_instance], evaluations=evaluations, num_challenge=4,
                            numerator=E.distribute_powers(constraints, alpha))


# ----------------------------------------------------------------------------------------------------------------------
# the multi-open verifier's MSM scalars as a program: `Msm` (util/msm.rs) over virtual registers
# ----------------------------------------------------------------------------------------------------------------------
class SymbolicMsm:
    """util/msm.rs:20-226 with ProgramBuilder values as scalars and BASE SLOTS (small integers) instead of curve points: the
    bases of a proof differ from proof to proof, the slot a base occupies in the final MSM does not.  `push` dedupes by slot, as
    the reference dedupes by point equality (msm.rs:109-116)."""

    def __init__(self, b: ProgramBuilder, constant=None, terms=None):
        self.b, self.constant, self.terms = b, constant, dict(terms or {})     # terms: slot -> scalar value (insertion ordered)

    @staticmethod
    def base(b, slot):
        return SymbolicMsm(b, None, {slot: b.const(1)})

    @staticmethod
    def constant_(b, value):
        return SymbolicMsm(b, value)

    def scale(self, k):
        c = None if self.constant is None else self.b.mul(self.constant, k)
        return SymbolicMsm(self.b, c, {s: self.b.mul(v, k) for s, v in self.terms.items()})

    def __mul__(self, k):
        return self.scale(k)

    def __neg__(self):
        return SymbolicMsm(self.b, None if self.constant is None else self.b.neg(self.constant), {s: self.b.neg(v) for s, v in self.terms.items()})

    def __add__(self, o):
        if self.constant is None:
            c = o.constant
        elif o.constant is None:
            c = self.constant
        else:
            c = self.b.add(self.constant, o.constant)
        t = dict(self.terms)
        for s, v in o.terms.items():
            t[s] = self.b.add(t[s], v) if s in t else v
        return SymbolicMsm(self.b, c, t)

    def __sub__(self, o):
        return self + (-o)

    @staticmethod
    def sum(b, msms):
        acc = SymbolicMsm(b)
        for m in msms:
            acc = acc + m
        return acc


def _powers(b: ProgramBuilder, x: int, n: int) -> List[int]:
    """loader.rs:71-78"""
    out = [b.const(1)]
    if n > 1:
        out.append(x)
    for _ in range(2, n):
        out.append(b.mul(out[-1], x))
    return out[:n]


@dataclass
class MsmScalarProgram:
    """`program` outputs, per proof, the scalars of the lhs MSM followed by those of the rhs MSM; `lhs_slots` / `rhs_slots` name
    the base behind each scalar: ("g",) = the SRS generator (carries the Msm constant, msm.rs:81-98), ("c", poly) = commitment of
    polynomial `poly`, ("w", i) = the i-th opening proof point."""
    program: Program
    lhs_slots: List[tuple]
    rhs_slots: List[tuple]
    input_layout: Dict[str, int]


def gwc19_symbolic(b: ProgramBuilder, commitments, z: int, queries: Sequence[Tuple[int, int, int]], v: int, u: int, w_slot=lambda i: ("w", i)):
    """`Gwc19::verify` (pcs/kzg/multiopen/gwc19.rs:45-82) over builder values.  `commitments`: list of SymbolicMsm (or a callable
    returning it — evaluated where the reference first touches the commitments, which fixes the instruction order); `queries`:
    (poly, shift constant, evaluation VALUE) in protocol order.  -> (lhs, rhs) SymbolicMsm"""
    sets = []                                                           # gwc19.rs:140-160
    for poly, shift, ev in queries:
        for st in sets:
            if st["shift"] == shift:
                st["polys"].append(poly); st["evals"].append(ev)
                break
        else:
            sets.append({"shift": shift, "polys": [poly], "evals": [ev]})
    powers_of_u = _powers(b, u, len(sets))
    powers_of_v = _powers(b, v, max(len(st["polys"]) for st in sets))
    if callable(commitments):
        commitments = commitments()
    f = SymbolicMsm(b)
    for st, pu in zip(sets, powers_of_u):
        set_msm = SymbolicMsm(b)
        for poly, ev, pv in zip(st["polys"], st["evals"], powers_of_v):
            set_msm = set_msm + (commitments[poly] - SymbolicMsm.constant_(b, ev)) * pv
        f = f + set_msm * pu
    z_omegas = [b.mul(b.const(st["shift"]), z) for st in sets]
    rhs = [SymbolicMsm.base(b, w_slot(i)) * pu for i, pu in enumerate(powers_of_u)]
    lhs = f + SymbolicMsm.sum(b, [uw * zo for uw, zo in zip(rhs, z_omegas)])
    return lhs, SymbolicMsm.sum(b, rhs)


def gwc19_num_sets(queries) -> int:
    """number of opening-proof points `Gwc19Proof::read` reads (gwc19.rs:103-107): one per distinct shift"""
    return len({q[1] for q in queries})


def compile_gwc19_msm_scalars(queries: Sequence[Tuple[int, int]], num_polys: int) -> MsmScalarProgram:
    """`Gwc19::verify` as a program for plain commitments.  `queries`: (poly, shift) in protocol order; every
    commitment is a plain base (`Msm::base`), which is what the halo2 system produces without linearization.  Per-proof input row:
    [z | v | u | one evaluation per query, in query order]."""
    b = ProgramBuilder()
    lay = {"z": 0, "v": 1, "u": 2, "evals": 3, "total": 3 + len(queries)}
    z, v, u = b.input(0), b.input(1), b.input(2)
    qs = [(poly, shift, b.input(lay["evals"] + k)) for k, (poly, shift) in enumerate(queries)]
    lhs, rhs_sum = gwc19_symbolic(b, lambda: [SymbolicMsm.base(b, ("c", j)) for j in range(num_polys)], z, qs, v, u)
    return _finish_msm_program(b, lhs, rhs_sum, lay)


def _finish_msm_program(b, lhs: SymbolicMsm, rhs: SymbolicMsm, lay) -> MsmScalarProgram:
    def flat(m):
        slots, vals = [], []
        if m.constant is not None:                                      # evaluate(Some(gen)): (constant, gen) goes first
            slots.append(("g",)); vals.append(m.constant)
        for s, v in m.terms.items():
            slots.append(s); vals.append(v)
        return slots, vals
    ls, lv = flat(lhs)
    rs, rv = flat(rhs)
    b.n_inputs = lay["total"]
    prog = b.finish(lv + rv, ["lhs%d" % i for i in range(len(lv))] + ["rhs%d" % i for i in range(len(rv))])
    return MsmScalarProgram(prog, ls, rs, lay)


def compile_bdfg21_msm_scalars(queries: Sequence[Tuple[int, int]], num_polys: int) -> MsmScalarProgram:
    """`Bdfg21::verify` as a program for plain commitments.  `queries`: (poly, shift) in protocol order.  Per-proof input row:
    [z | mu | gamma | z' | one evaluation per query, in query order].  Opening-proof slots: ("w", 0) = W, ("w", 1) = W'."""
    b = ProgramBuilder()
    lay = {"z": 0, "mu": 1, "gamma": 2, "z_prime": 3, "evals": 4, "total": 4 + len(queries)}
    z, mu, gamma, z_prime = (b.input(i) for i in range(4))
    qs = [(poly, shift, b.input(lay["evals"] + k)) for k, (poly, shift) in enumerate(queries)]
    lhs, rhs = bdfg21_symbolic(b, lambda: [SymbolicMsm.base(b, ("c", j)) for j in range(num_polys)], z, qs, mu, gamma, z_prime)
    return _finish_msm_program(b, lhs, rhs, lay)


def bdfg21_symbolic(b: ProgramBuilder, commitments, z: int, queries: Sequence[Tuple[int, int, int]], mu: int, gamma: int, z_prime: int,
                    w_slot=lambda i: ("w", i)):
    """`Bdfg21::verify` (pcs/kzg/multiopen/bdfg21.rs:51-83; query sets :123-175, coefficients :177-371) over builder values.
    `queries`: (poly, shift constant, evaluation VALUE) in protocol order.  The shift arithmetic (normalised ell', set grouping)
    depends only on the protocol and is done here on the host; everything that depends on z, z', mu, gamma or the evaluations
    becomes instructions, with the two rounds of `L::batch_invert` (bdfg21.rs:218-219) as two shared inversions.
    -> (lhs, rhs) SymbolicMsm"""
    R = R_MODULUS
    # ---- query_sets (bdfg21.rs:123-175) on (poly, shift, eval register) ----
    poly_shifts = []
    for poly, shift, ev in queries:
        for ps in poly_shifts:
            if ps[0] == poly:
                if shift not in ps[1]:
                    ps[1].append(shift); ps[2].append(ev)
                break
        else:
            poly_shifts.append((poly, [shift], [ev]))
    sets = []
    for poly, shifts, evals in poly_shifts:
        for st in sets:
            if set(st["shifts"]) == set(shifts):
                if poly not in st["polys"]:
                    st["polys"].append(poly)
                    st["evals"].append([evals[shifts.index(lhs)] for lhs in st["shifts"]])
                break
        else:
            sets.append({"shifts": shifts, "polys": [poly], "evals": [evals]})
    # ---- query_set_coeffs (bdfg21.rs:177-222) ----
    superset = sorted({s for st in sets for s in st["shifts"]})
    size = max([len(st["shifts"]) for st in sets] + [2])
    powers_of_z = _powers(b, z, size)
    zp_minus = {s: b.sub(z_prime, b.mul(z, b.const(s))) for s in superset}
    coeffs, z_s_1 = [], None
    for st in sets:
        shifts = st["shifts"]
        ell = []
        for j, sj in enumerate(shifts):
            acc = 1
            for i, si in enumerate(shifts):
                if i != j:
                    acc = acc * (sj - si) % R
            ell.append(acc)
        zz, z_pow = powers_of_z[1], powers_of_z[len(shifts) - 1]
        bary = []
        for s, e in zip(shifts, ell):                                   # sum_products_with_coeff (bdfg21.rs:297-302)
            t1 = b.mul(b.mul(z_pow, z_prime), b.const(e))
            t2 = b.mul(b.mul(z_pow, zz), b.const((-(e * s)) % R))
            bary.append(b.add(t1, t2))
        z_s = None
        for s in shifts:                                                # loader.product (bdfg21.rs:307-312)
            z_s = zp_minus[s] if z_s is None else b.mul(z_s, zp_minus[s])
        coeffs.append({"bary": bary, "z_s": z_s, "z_s_1": z_s_1})
        if z_s_1 is None:
            z_s_1 = z_s
    # first batch_invert: barycentric weights (+ z_s of every set but the first)
    den1 = []
    for c in coeffs:
        den1 += c["bary"] + ([c["z_s"]] if c["z_s_1"] is not None else [])
    inv1 = b.batch_invert(den1)
    k = 0
    den2 = []
    for c in coeffs:
        n = len(c["bary"])
        c["eval_coeffs"] = inv1[k:k + n]                                # Fraction::one_over(..).evaluate()
        k += n
        if c["z_s_1"] is not None:
            c["commitment_coeff"] = b.mul(c["z_s_1"], inv1[k])          # Fraction::new(z_s_1, z_s).evaluate()
            k += 1
        else:
            c["commitment_coeff"] = None
        wsum = None
        for w in c["eval_coeffs"]:                                      # loader.sum (bdfg21.rs:345-351)
            wsum = w if wsum is None else b.add(wsum, w)
        den2.append(wsum)
    inv2 = b.batch_invert(den2)                                         # second batch_invert: the r_eval coefficients
    for c, iv in zip(coeffs, inv2):
        c["r_eval_coeff"] = iv if c["commitment_coeff"] is None else b.mul(c["commitment_coeff"], iv)
    # ---- verify (bdfg21.rs:58-82) ----
    powers_of_mu = _powers(b, mu, max(len(st["polys"]) for st in sets))
    powers_of_gamma = _powers(b, gamma, len(sets))
    if callable(commitments):
        commitments = commitments()
    f = SymbolicMsm(b)
    for st, co, pg in zip(sets, coeffs, powers_of_gamma):
        set_msm = SymbolicMsm(b)
        for poly, evals, pm in zip(st["polys"], st["evals"], powers_of_mu):
            commitment = commitments[poly] * co["commitment_coeff"] if co["commitment_coeff"] is not None else commitments[poly]
            r_eval = None
            for cf, ev in zip(co["eval_coeffs"], evals):               # loader.sum_products
                t = b.mul(cf, ev)
                r_eval = t if r_eval is None else b.add(r_eval, t)
            r_eval = b.mul(r_eval, co["r_eval_coeff"])
            set_msm = set_msm + (commitment - SymbolicMsm.constant_(b, r_eval)) * pm
        f = f + set_msm * pg
    f = f - SymbolicMsm.base(b, w_slot(0)) * coeffs[0]["z_s"]
    rhs = SymbolicMsm.base(b, w_slot(1))
    lhs = f + rhs * z_prime
    return lhs, rhs
