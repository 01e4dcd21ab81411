"""`PlonkProtocol`, `PlonkProof`, `PlonkSuccinctVerifier` / `PlonkVerifier` driven by a protocol object.

Mirrored reference items (paths relative to snark-verifier/src):
  verifier/plonk/protocol.rs:19-106     PlonkProtocol (fields, `langranges`), :285-300 QuotientPolynomial, :465-470 LinearizationStrategy
  verifier/plonk/proof.rs:52-169        PlonkProof::read — the transcript schedule: initial state, instances, witnesses per phase with
                                        their challenges, quotient chunks, z, evaluations, the multi-open proof, old accumulators
  verifier/plonk/proof.rs:171-199       empty_queries / queries (shift = omega^rotation)
  verifier/plonk/proof.rs:201-306       commitments: bases, the linearized numerator as an Msm, quotient = sum (z^n)^(chunk_degree i) t_i,
                                        the three linearization strategies
  verifier/plonk/proof.rs:308-349       evaluations (instance evaluations by Lagrange, then the proof's own)
  verifier/plonk.rs:57-93, 113-134      PlonkSuccinctVerifier::verify -> accumulators; PlonkVerifier::verify = + decide_all
  system/halo2/transcript/evm.rs:175-268  EvmTranscript over NativeLoader: 32-byte big-endian words, `from_xy` / canonical checks

How it runs here.  The protocol of a batch is fixed, so everything the reference computes per proof in `Fr` is compiled ONCE into a
straight-line program (`compile_plonk_verifier`): the reference code paths above are executed with `plonk_eval.ProgramBuilder`
values as scalars and `plonk_eval.SymbolicMsm` as `Msm` — exactly how the reference drives its EVM / Halo2 loaders — and end in the
multi-open verifier (`gwc19_symbolic` / `bdfg21_symbolic`).  The program's outputs are the scalars of the two final MSMs.  Per batch:
  Keccak transcript challenges of all proofs   -> snarkv_evm_transcript_challenges   (one thread per proof)
  per-proof MSM scalars                        -> snarkv_fr_program_eval_batch       (one thread per proof)
  sum_j rho^j lhs_j, sum_j rho^j rhs_j          -> snarkv_g1_msm_batch_rlc x 2        (every proof point validated on the device)
  e(lhs, g2) e(rhs, -s g2) == 1                 -> snarkv_kzg_decide_batch
A single proof is the batch of one (rho = 1), which is the reference's `PlonkVerifier::verify`.
"""
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import CHECK_INPUTS, AssertionFailure, Error, KzgAccumulator
from .pcs import LimbsEncoding
from .plonk_eval import (R_MODULUS, CommonPolynomial, CommonPolynomialEvaluation, Domain, Expression, MsmScalarProgram, ProgramBuilder,
                         Query, Rotation, SymbolicMsm, _finish_msm_program, bdfg21_symbolic, gwc19_num_sets, gwc19_symbolic)

P_MODULUS = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47

WITHOUT_CONSTANT = "WithoutConstant"                                    # protocol.rs:465-470
MINUS_VANISHING_TIMES_QUOTIENT = "MinusVanishingTimesQuotient"


class InvalidInstances(Error):
    """Error::InvalidInstances (lib.rs:18-28)"""


class InvalidProtocol(Error):
    """Error::InvalidProtocol(String)"""


class TranscriptError(Error):
    """Error::Transcript(io::ErrorKind::InvalidData, String)"""


@dataclass
class QuotientPolynomial:
    """protocol.rs:285-300"""
    chunk_degree: int
    num_chunk: int
    numerator: Expression


@dataclass
class PlonkProtocol:
    """protocol.rs:19-67.  `preprocessed`: 64-byte affine points (canonical little-endian x || y)."""
    domain: Domain
    preprocessed: List[bytes]
    num_instance: List[int]
    num_witness: List[int]
    num_challenge: List[int]
    evaluations: List[Query]
    queries: List[Query]
    quotient: QuotientPolynomial
    transcript_initial_state: Optional[int] = None
    linearization: Optional[str] = None
    accumulator_indices: List[List[Tuple[int, int]]] = field(default_factory=list)
    # instance_committing_key: not supported (the halo2 KZG system never sets it; proof.rs:75-101)

    def langranges(self) -> List[int]:
        """protocol.rs:77-106: the numerator's own Lagrange indices plus the range the instance evaluations need"""
        offset = len(self.preprocessed)
        rng = range(offset, offset + len(self.num_instance))
        min_rot = max_rot = 0
        for q in sorted(self.quotient.numerator.used_query()):
            if q.poly in rng:
                if q.rotation.value < min_rot:
                    min_rot = q.rotation.value
                elif q.rotation.value > max_rot:
                    max_rot = q.rotation.value
        max_instance_len = max(self.num_instance) if self.num_instance else 0
        return sorted(set(self.quotient.numerator.used_langrange()) | set(range(-max_rot, max_instance_len + abs(min_rot))))

    def num_polys(self) -> int:
        """preprocessed, instance, witness polynomials, then the quotient (+ the linearization polynomial)"""
        return len(self.preprocessed) + len(self.num_instance) + sum(self.num_witness) + (2 if self.linearization == WITHOUT_CONSTANT else 1)


# ----------------------------------------------------------------------------------------------------------------------
# a small hand-built protocol (tests / bench): what system/halo2.rs would describe for a width-3 arithmetisation without a
# permutation argument — fixed q_l, q_r, q_o, q_m, q_c, q_next (0..5), one instance column (6), advice a, b, c (7..9), quotient (10):
#   gate:  q_l a + q_r b + q_o c + q_m a b + q_c + instance = 0        next:  q_next (a(wX) - c(X)) = 0
# `tests/plonk_toy.py` is the from-scratch prover for it.
# ----------------------------------------------------------------------------------------------------------------------
def simple_plonk_protocol(k: int, preprocessed: Sequence[bytes], num_public: int, linearization: Optional[str] = None,
                          initial_state: Optional[int] = None, accumulator_indices=()) -> PlonkProtocol:
    E = Expression
    poly = lambda i, r=0: E.polynomial(Query(i, Rotation(r)))
    ql, qr, qo, qm, qc, qn = (poly(i) for i in range(6))
    inst, a, b, c, a_next = poly(6), poly(7), poly(8), poly(9), poly(7, 1)
    # the evaluated part enters with a minus sign when the verifier keeps `constant` itself as the evaluation (proof.rs:283-290)
    gate = ql * a + qr * b + qo * c + qm * a * b + qc + (-inst if linearization == MINUS_VANISHING_TIMES_QUOTIENT else inst)
    numerator = E.distribute_powers([gate, qn * (a_next - c)], E.challenge(0))
    fixed_eval = [Query(i, Rotation(0)) for i in range(6)]
    advice_eval = [Query(7, Rotation(0)), Query(7, Rotation(1)), Query(8, Rotation(0)), Query(9, Rotation(0))]
    quotient_q = Query(10, Rotation(0))
    if linearization is None:                                   # every polynomial evaluated; quotient evaluation computed by the verifier
        evaluations = fixed_eval + advice_eval
        queries = evaluations + [quotient_q]
    elif linearization == WITHOUT_CONSTANT:                     # selectors stay commitments; the proof carries r(z) of the linearization polynomial
        lin_q = Query(11, Rotation(0))
        evaluations = advice_eval + [lin_q]
        queries = advice_eval + [quotient_q, lin_q]
    else:
        evaluations = advice_eval
        queries = advice_eval + [quotient_q]
    return PlonkProtocol(domain=Domain(k), preprocessed=[bytes(p) for p in preprocessed], num_instance=[num_public], num_witness=[3],
                         num_challenge=[1], evaluations=evaluations, queries=queries, quotient=QuotientPolynomial(1, 2, numerator),
                         transcript_initial_state=initial_state, linearization=linearization,
                         accumulator_indices=[list(x) for x in accumulator_indices])


# ----------------------------------------------------------------------------------------------------------------------
# transcript schedule of PlonkProof::read for the Keccak EvmTranscript: everything absorbed is a 32-byte big-endian word
# ----------------------------------------------------------------------------------------------------------------------
class TranscriptLayout:
    """Word offsets (32-byte words) inside the absorbed stream  [initial state | instances | proof bytes]  and the squeeze points,
    in the order PlonkProof::read (proof.rs:52-169) and the multi-open `read_proof` (gwc19.rs:100-108, bdfg21.rs:105-118) produce them."""

    def __init__(self, protocol: PlonkProtocol, scheme: str):
        assert scheme in ("gwc19", "bdfg21")
        self.scheme = scheme
        pos = 0
        self.initial_state = None
        if protocol.transcript_initial_state is not None:
            self.initial_state = pos
            pos += 1
        self.instances = pos
        pos += sum(protocol.num_instance)
        self.proof_start = pos
        self.seg_end: List[int] = []                                    # word offset after which each challenge is squeezed
        self.witnesses: List[int] = []
        for n, m in zip(protocol.num_witness, protocol.num_challenge):  # proof.rs:113-135
            for _ in range(n):
                self.witnesses.append(pos)
                pos += 2
            self.seg_end += [pos] * m
        self.n_challenges = len(self.seg_end)
        self.quotients = [pos + 2 * i for i in range(protocol.quotient.num_chunk)]
        pos += 2 * protocol.quotient.num_chunk
        self.seg_end.append(pos)                                        # z
        self.evaluations = pos
        pos += len(protocol.evaluations)
        shifts = [protocol.domain.rotate_scalar(1, q.rotation) for q in protocol.queries]
        if scheme == "gwc19":
            self.seg_end.append(pos)                                    # v
            n_sets = gwc19_num_sets([(q.poly, s) for q, s in zip(protocol.queries, shifts)])
            self.ws = [pos + 2 * i for i in range(n_sets)]
            pos += 2 * n_sets
            self.seg_end.append(pos)                                    # u
        else:
            self.seg_end += [pos, pos]                                  # mu, gamma
            self.ws = [pos]                                             # W
            pos += 2
            self.seg_end.append(pos)                                    # z'
            self.ws.append(pos)                                         # W'
            pos += 2
        self.total = pos
        self.proof_words = pos - self.proof_start
        # the proof's own items in stream order: (kind, word offset).  EvmTranscript serialises a point as two 32-byte big-endian
        # words; the Poseidon transcript reads 32-byte compressed points and 32-byte little-endian scalars (halo2.rs:244-274) but
        # ABSORBS the same count of field elements per item, so the word offsets double as element offsets.
        self.items = sorted([("point", o) for o in self.witnesses + self.quotients + self.ws]
                            + [("scalar", self.evaluations + i) for i in range(len(protocol.evaluations))], key=lambda it: it[1])

    def proof_len(self, transcript: str = "evm") -> int:
        return 32 * (self.proof_words if transcript == "evm" else len(self.items))


def _from_xy(x: int, y: int) -> bytes:
    """EvmTranscript::read_ec_point (transcript/evm.rs:247-266): canonical coordinates on the curve, else InvalidData.  (0, 0) is not
    on the curve, so — as in the reference's `from_xy` — a proof cannot carry the identity."""
    if x >= P_MODULUS or y >= P_MODULUS or (y * y - x * x * x - 3) % P_MODULUS != 0:
        raise TranscriptError("Invalid elliptic curve point encoding in proof")
    return x.to_bytes(32, "little") + y.to_bytes(32, "little")


# ----------------------------------------------------------------------------------------------------------------------
# the verifier's Fr half as ONE program: PlonkProof::{evaluations, commitments, queries} + the multi-open verifier
# ----------------------------------------------------------------------------------------------------------------------
@dataclass
class PlonkVerifierProgram:
    msm: MsmScalarProgram            # program + base slots of the lhs / rhs MSM
    layout: Dict[str, int]           # per-proof input row: [instances | challenges | z | evaluations | multi-open challenges]
    transcript: TranscriptLayout


def compile_plonk_verifier(protocol: PlonkProtocol, scheme: str = "gwc19") -> PlonkVerifierProgram:
    """verifier/plonk.rs:57-93 over builder values.  Base slots of the result: ("g",) the SRS generator, ("pre", i) preprocessed
    commitment i (shared by all proofs), ("wit", i) witness commitment, ("quot", i) quotient chunk, ("w", i) opening-proof point."""
    tl = TranscriptLayout(protocol, scheme)
    b = ProgramBuilder()
    n_inst, n_chal, n_eval = sum(protocol.num_instance), tl.n_challenges, len(protocol.evaluations)
    n_pcs = 2 if scheme == "gwc19" else 3
    lay = {"instances": 0, "challenges": n_inst, "z": n_inst + n_chal, "evaluations": n_inst + n_chal + 1}
    lay["pcs"] = lay["evaluations"] + n_eval
    lay["total"] = lay["pcs"] + n_pcs
    z = b.input(lay["z"])
    challenges = [b.input(lay["challenges"] + i) for i in range(n_chal)]
    inst_base = [lay["instances"] + sum(protocol.num_instance[:k]) for k in range(len(protocol.num_instance))]

    # CommonPolynomialEvaluation::new + batch_invert(denoms) + evaluate (plonk.rs:63-72)
    cpe = CommonPolynomialEvaluation(b, protocol.domain, protocol.langranges(), z)

    # PlonkProof::evaluations (proof.rs:308-349)
    offset = len(protocol.preprocessed)
    inst_range = range(offset, offset + len(protocol.num_instance))
    evaluations: Dict[Query, int] = {}
    for q in sorted(qq for qq in protocol.quotient.numerator.used_query() if qq.poly in inst_range):
        col = q.poly - offset
        acc = None
        for j in range(protocol.num_instance[col]):                     # loader.sum_products(instances x L_{j - rotation})
            term = b.mul(b.input(inst_base[col] + j), cpe.get(CommonPolynomial.lagrange(j - q.rotation.value)))
            acc = term if acc is None else b.add(acc, term)
        evaluations[q] = acc if acc is not None else b.const(0)
    for k, q in enumerate(protocol.evaluations):
        evaluations[q] = b.input(lay["evaluations"] + k)

    # PlonkProof::commitments (proof.rs:201-306)
    n_wit = sum(protocol.num_witness)
    commitments: List[SymbolicMsm] = [SymbolicMsm.base(b, ("pre", i)) for i in range(len(protocol.preprocessed))]
    commitments += [SymbolicMsm(b) for _ in protocol.num_instance]     # no instance committing key: Default (empty) Msm
    commitments += [SymbolicMsm.base(b, ("wit", i)) for i in range(n_wit)]

    def msm_query(q: Query) -> SymbolicMsm:
        if q in evaluations:
            return SymbolicMsm.constant_(b, evaluations[q])
        if q.rotation.value == 0 and q.poly < len(commitments):
            return commitments[q.poly]
        raise InvalidProtocol("Missing query %r" % (q,))

    def msm_challenge(i: int) -> SymbolicMsm:
        if i >= len(challenges):
            raise InvalidProtocol("Missing challenge %d" % i)
        return SymbolicMsm.constant_(b, challenges[i])

    def msm_product(x: SymbolicMsm, y: SymbolicMsm) -> SymbolicMsm:
        if not x.terms:                                                 # (0, _) => b * a.try_into_constant()
            return y * _constant_of(b, x)
        if not y.terms:
            return x * _constant_of(b, y)
        raise InvalidProtocol("Invalid linearization")

    numerator = protocol.quotient.numerator.evaluate(
        lambda c: SymbolicMsm.constant_(b, b.const(c)), lambda p: SymbolicMsm.constant_(b, cpe.get(p)), msm_query, msm_challenge,
        lambda x: -x, lambda x, y: x + y, msm_product, lambda x, c: x * b.const(c))

    quotient_query = Query(len(protocol.preprocessed) + len(protocol.num_instance) + n_wit, Rotation(0))
    zn_pow = cpe.zn if protocol.quotient.chunk_degree == 1 else b.pow_const(cpe.zn, protocol.quotient.chunk_degree)
    quotient = SymbolicMsm(b)
    coeff = None
    for i in range(protocol.quotient.num_chunk):                        # powers(num_chunk) zipped with the chunks
        coeff = b.const(1) if i == 0 else (zn_pow if i == 1 else b.mul(coeff, zn_pow))
        quotient = quotient + SymbolicMsm.base(b, ("quot", i)) * coeff
    if protocol.linearization == WITHOUT_CONSTANT:
        linearization_query = Query(quotient_query.poly + 1, Rotation(0))
        constant = numerator.constant if numerator.constant is not None else b.const(0)
        commitments.append(quotient)
        commitments.append(SymbolicMsm(b, None, numerator.terms))
        evaluations[quotient_query] = b.mul(b.add(constant, evaluations[linearization_query]), cpe.zn_minus_one_inv)
    elif protocol.linearization == MINUS_VANISHING_TIMES_QUOTIENT:
        lin = numerator - quotient * cpe.zn_minus_one
        commitments.append(SymbolicMsm(b, None, lin.terms))
        evaluations[quotient_query] = lin.constant if lin.constant is not None else b.const(0)
    else:
        if numerator.terms:
            raise InvalidProtocol("Invalid linearization")
        commitments.append(quotient)
        evaluations[quotient_query] = b.mul(_constant_of(b, numerator), cpe.zn_minus_one_inv)

    # PlonkProof::queries (proof.rs:171-199): shift = omega^rotation, evaluation looked up by Query
    queries = []
    for q in protocol.queries:
        if q not in evaluations:
            raise InvalidProtocol("Missing evaluation of %r" % (q,))
        queries.append((q.poly, protocol.domain.rotate_scalar(1, q.rotation), evaluations[q]))

    pcs = [b.input(lay["pcs"] + i) for i in range(n_pcs)]
    if scheme == "gwc19":
        lhs, rhs = gwc19_symbolic(b, commitments, z, queries, pcs[0], pcs[1])
    else:
        lhs, rhs = bdfg21_symbolic(b, commitments, z, queries, pcs[0], pcs[1], pcs[2])
    return PlonkVerifierProgram(_finish_msm_program(b, lhs, rhs, lay), lay, tl)


def _constant_of(b: ProgramBuilder, m: SymbolicMsm) -> int:
    """Msm::try_into_constant (util/msm.rs:67-70) for a base-free Msm; an empty Msm has no constant: the reference unwraps"""
    if m.terms or m.constant is None:
        raise InvalidProtocol("Invalid linearization")
    return m.constant


# ----------------------------------------------------------------------------------------------------------------------
# PlonkProof / verifiers
# ----------------------------------------------------------------------------------------------------------------------
@dataclass
class PlonkProof:
    """proof.rs:20-44 for one proof (host values: points as 64-byte canonical LE, scalars as ints)."""
    witnesses: List[bytes]
    challenges: List[int]
    quotients: List[bytes]
    z: int
    evaluations: List[int]
    pcs_challenges: List[int]        # gwc19: [v, u]; bdfg21: [mu, gamma, z']
    pcs_points: List[bytes]          # gwc19: ws; bdfg21: [W, W']
    old_accumulators: List[KzgAccumulator]


class PlonkBatchVerifier:
    """`PlonkVerifier<KzgAs<Bn256, Gwc19 | Bdfg21>, LimbsEncoding<LIMBS, BITS>>` (verifier/plonk.rs:96-134) for m proofs of ONE protocol,
    the m accumulators (and their old accumulators) fused by a random linear combination (pcs/kzg/decider.rs:146-185) into one
    pairing check.  `kzg`: a `KzgAs` holding the deciding key; `svk_g`: the SRS generator."""

    def __init__(self, loader, kzg, svk_g: bytes, protocol: PlonkProtocol, scheme: str = "gwc19", limbs: Tuple[int, int] = (4, 68),
                 transcript: str = "evm"):
        # transcript: Keccak `EvmTranscript` (transcript/evm.rs) | "poseidon" = PoseidonTranscript<_, _, _, 5, 4, 8, 60> (transcript/halo2.rs)
        assert transcript in ("evm", "poseidon")
        self.loader, self.kzg, self.svk_g, self.protocol, self.scheme, self.transcript = loader, kzg, bytes(svk_g), protocol, scheme, transcript
        self.compiled = compile_plonk_verifier(protocol, scheme)
        self.tl = self.compiled.transcript
        self.encoding = LimbsEncoding(*limbs)
        self._slots = {"lhs": self.compiled.msm.lhs_slots, "rhs": self.compiled.msm.rhs_slots}
        self._plan = None
        self.use_device_plan = hasattr(loader, "plonk_plan_create")   # the fused device-resident pipeline (csrc/plonk_batch.cu)

    def close(self):
        if self._plan is not None:
            self.loader.plonk_plan_free(self._plan)
            self._plan = None

    # -- the whole batch in ONE device call (csrc/plonk_batch.cu): streams up, fused accumulator (and verdict) down ------------------------
    def _device_plan(self):
        if self._plan is None:
            tl, lay, pr = self.tl, self.compiled.layout, self.protocol
            n_inst, n_chal, n_eval = lay["challenges"], tl.n_challenges, len(pr.evaluations)
            row_src = [tl.instances + i for i in range(n_inst)] + [-(c + 1) for c in range(n_chal + 1)]          # instances | challenges, z
            row_src += [tl.evaluations + i for i in range(n_eval)] + [-(c + 1) for c in range(n_chal + 1, len(tl.seg_end))]
            row_check = [0] * (n_inst + n_chal + 1) + [1] * n_eval + [0] * (len(tl.seg_end) - n_chal - 1)          # read_scalar: evaluations < r
            consts, index = [], {}

            def const_slot(pt):
                if pt not in index:
                    index[pt] = len(consts)
                    consts.append(pt)
                return -(index[pt] + 1)
            poseidon = None
            point_src = {o: o for o in tl.witnesses + tl.quotients + tl.ws}          # EvmTranscript: a point is addressed by its word offset
            if self.transcript == "poseidon":
                pt_offsets = [off for kind, off in tl.items if kind == "point"]
                point_src = {off: k for k, off in enumerate(pt_offsets)}             # Poseidon: by its index among the decompressed points
                item_off = [off if kind == "scalar" else -(off + 1) for kind, off in tl.items]
                item_pt = [point_src[off] if kind == "point" else -1 for kind, off in tl.items]
                poseidon = (tl.proof_start, item_off, item_pt, len(pt_offsets))
            src = {}
            for side in ("lhs", "rhs"):
                src[side] = []
                for s in self._slots[side]:
                    if s == ("g",):
                        src[side].append(const_slot(self.svk_g))
                    elif s[0] == "pre":
                        src[side].append(const_slot(pr.preprocessed[s[1]]))
                    else:
                        src[side].append(point_src[{"wit": tl.witnesses, "quot": tl.quotients, "w": tl.ws}[s[0]][s[1]]])
            self._plan = self.loader.plonk_plan_create(tl.total, list(tl.seg_end), self.compiled.msm.program, row_src, row_check, src["lhs"], src["rhs"], consts,
                                                       poseidon)
        return self._plan

    def _streams_wire(self, instances, proofs) -> np.ndarray:
        """m x bytes handed to the device plan: [initial state | instances | proof].  EvmTranscript: 32-byte big-endian words (this IS the
        absorbed stream); Poseidon: little-endian scalars followed by the proof's 32-byte items (compressed points / little-endian scalars)"""
        tl, pr = self.tl, self.protocol
        order = "big" if self.transcript == "evm" else "little"
        m, plen, shape = len(proofs), tl.proof_len(self.transcript), list(pr.num_instance)
        if isinstance(proofs, np.ndarray):
            if proofs.shape != (m, plen):
                raise TranscriptError("proofs: %r, the protocol's transcript reads %d bytes per proof" % (proofs.shape, plen))
        else:
            for j, proof in enumerate(proofs):
                if len(proof) != plen:
                    raise TranscriptError("proof %d: %d bytes, the protocol's transcript reads %d" % (j, len(proof), plen))
            proofs = np.frombuffer(b"".join(proofs), dtype=np.uint8).reshape(m, plen)
        st = np.empty((m, 32 * tl.proof_start + plen), dtype=np.uint8)
        if tl.initial_state is not None:
            st[:, :32] = np.frombuffer((pr.transcript_initial_state % R_MODULUS).to_bytes(32, order), dtype=np.uint8)
        n_inst = sum(shape)
        if isinstance(instances, np.ndarray):                     # m x n_inst x 32 B words in the transcript's byte order, packed by the caller
            if instances.shape != (m, n_inst, 32):
                raise InvalidInstances("instances: %r, expected %r" % (instances.shape, (m, n_inst, 32)))
            st[:, 32 * tl.instances:32 * tl.proof_start] = instances.reshape(m, -1)
        else:
            for j, inst in enumerate(instances):
                if [len(col) for col in inst] != shape:
                    raise InvalidInstances("proof %d: instance column lengths %r != %r" % (j, [len(c) for c in inst], pr.num_instance))
            if n_inst:
                words = b"".join((v % R_MODULUS).to_bytes(32, order) for inst in instances for col in inst for v in col)
                st[:, 32 * tl.instances:32 * tl.proof_start] = np.frombuffer(words, dtype=np.uint8).reshape(m, 32 * n_inst)
        st[:, 32 * tl.proof_start:] = proofs
        return st

    def _accumulate_new_device(self, instances, proofs, rho: int, decide: bool):
        st = self._streams_wire(instances, proofs)
        try:
            lhs, rhs, ok = self.loader.plonk_accumulate_batch(self._device_plan(), st, st.shape[0], (rho % R_MODULUS).to_bytes(32, "little"), decide)
        except Error as e:
            if "Invalid" in str(e):                                # Error::Transcript(InvalidData, ..): a scalar >= r or a point off the curve
                raise TranscriptError(str(e)) from None
            raise
        return KzgAccumulator(lhs, rhs), ok

    # -- PlonkProof::read for a batch (proof.rs:52-169) -------------------------------------------------------------------
    def _parse(self, instances: Sequence[Sequence[Sequence[int]]], proofs: Sequence[bytes]):
        """-> (stream handed to the transcript kernel, m x total little-endian words, point lookup: word offset -> m x 64 B affine LE).
        EvmTranscript: the stream is [initial state | instances | proof] as 32-byte big-endian words, points are x || y in the proof.
        Poseidon: the stream is the absorbed field ELEMENTS; the proof's compressed points are decompressed and validated on the
        device (`C::from_bytes`), which also yields the two elements each point absorbs."""
        tl, pr, tr = self.tl, self.protocol, self.transcript
        m = len(proofs)
        plen, shape = tl.proof_len(tr), list(pr.num_instance)
        for j, (inst, proof) in enumerate(zip(instances, proofs)):
            if [len(col) for col in inst] != shape:
                raise InvalidInstances("proof %d: instance column lengths %r != %r" % (j, [len(c) for c in inst], pr.num_instance))
            if len(proof) != plen:
                raise TranscriptError("proof %d: %d bytes, the protocol's transcript reads %d" % (j, len(proof), plen))
        order = "big" if tr == "evm" else "little"
        st = np.zeros((m, tl.total, 32), dtype=np.uint8)
        if tl.initial_state is not None:
            st[:, 0] = np.frombuffer((pr.transcript_initial_state % R_MODULUS).to_bytes(32, order), dtype=np.uint8)
        n_inst = sum(shape)
        if n_inst:
            words = b"".join((v % R_MODULUS).to_bytes(32, order) for inst in instances for col in inst for v in col)
            st[:, tl.instances:tl.proof_start] = np.frombuffer(words, dtype=np.uint8).reshape(m, n_inst, 32)
        raw = np.frombuffer(b"".join(proofs), dtype=np.uint8).reshape(m, plen // 32, 32)
        if tr == "evm":
            st[:, tl.proof_start:] = raw
            words_le = st[:, :, ::-1]
            return st.reshape(m, -1), words_le, lambda w: np.concatenate([words_le[:, w], words_le[:, w + 1]], axis=1)
        pt_items = [i for i, (kind, _) in enumerate(tl.items) if kind == "point"]
        pts, els, valid = self.loader.g1_decompress(np.ascontiguousarray(raw[:, pt_items]).tobytes(), m * len(pt_items))
        if any(v != 1 for v in valid):
            raise TranscriptError("Invalid elliptic curve point encoding in proof")            # halo2.rs:263-268
        pts = np.frombuffer(pts, dtype=np.uint8).reshape(m, len(pt_items), 64)
        els = np.frombuffer(els, dtype=np.uint8).reshape(m, len(pt_items), 2, 32)
        by_off = {}
        for k, i in enumerate(pt_items):
            off = tl.items[i][1]
            st[:, off] = els[:, k, 0]
            st[:, off + 1] = els[:, k, 1]
            by_off[off] = pts[:, k]
        for i, (kind, off) in enumerate(tl.items):
            if kind == "scalar":
                st[:, off] = raw[:, i]
        return st.reshape(m, -1), st, lambda w: by_off[w]

    def read_proofs(self, instances, proofs):
        """PlonkProof::read for the batch -> (input rows m x n_inputs x 32 B LE, point lookup, challenges m x k x 32 B LE).  Scalars
        the proofs carry must be canonical (`read_scalar`); points are validated on the device — by the decompression kernel
        (Poseidon transcript) or by the MSM's input check (`from_xy`, EvmTranscript)."""
        tl, lay, m = self.tl, self.compiled.layout, len(proofs)
        st, words, lookup = self._parse(instances, proofs)
        if self.transcript == "evm":
            ch = self.loader.evm_transcript_challenges(st.tobytes(), tl.total * 32, [32 * e for e in tl.seg_end], m)
        else:
            ch = self.loader.poseidon_transcript_challenges(st.tobytes(), tl.total, list(tl.seg_end), m)
        ch = np.frombuffer(ch, dtype=np.uint8).reshape(m, len(tl.seg_end), 32)
        n_inst, n_chal, n_eval = lay["challenges"], tl.n_challenges, len(self.protocol.evaluations)
        rows = np.zeros((m, lay["total"], 32), dtype=np.uint8)
        rows[:, :n_inst] = words[:, tl.instances:tl.instances + n_inst]
        rows[:, lay["challenges"]:lay["challenges"] + n_chal + 1] = ch[:, :n_chal + 1]           # challenges, then z
        ev = words[:, tl.evaluations:tl.evaluations + n_eval]
        self._require_canonical(ev, "evaluation")
        rows[:, lay["evaluations"]:lay["evaluations"] + n_eval] = ev
        rows[:, lay["pcs"]:] = ch[:, n_chal + 1:]
        return rows, lookup, ch

    @staticmethod
    def _require_canonical(words_le: np.ndarray, what: str):
        """`read_scalar` (transcript/evm.rs:230-245): a scalar the proof carries must be < r"""
        flat = words_le.reshape(-1, 32)
        r = R_MODULUS.to_bytes(32, "little")
        suspects = flat[flat[:, 31] >= r[31]]                            # everything with a smaller top byte is < r
        for row in suspects:
            if int.from_bytes(row.tobytes(), "little") >= R_MODULUS:
                raise TranscriptError("Invalid scalar encoding in proof (%s >= r)" % what)

    def _points(self, lookup, m: int, side: str) -> np.ndarray:
        """m x n_slots x 64 B little-endian affine points for the slots of one MSM side"""
        tl = self.tl
        slots = self._slots[side]
        out = np.zeros((m, len(slots), 64), dtype=np.uint8)
        for k, s in enumerate(slots):
            if s == ("g",):
                out[:, k] = np.frombuffer(self.svk_g, dtype=np.uint8)
            elif s[0] == "pre":
                out[:, k] = np.frombuffer(self.protocol.preprocessed[s[1]], dtype=np.uint8)
            else:
                out[:, k] = lookup({"wit": tl.witnesses, "quot": tl.quotients, "w": tl.ws}[s[0]][s[1]])
        return out

    def old_accumulators(self, instances) -> List[KzgAccumulator]:
        """proof.rs:148-158: `AE::from_repr` over the instance values `accumulator_indices` points at"""
        accs = []
        for inst in instances:
            for idx in self.protocol.accumulator_indices:
                accs.append(self.encoding.from_repr([inst[i][j] for i, j in idx]))
        return accs

    # -- PlonkSuccinctVerifier::verify (plonk.rs:57-93) for the batch -------------------------------------------------------
    def accumulate_new(self, instances, proofs, rho: int) -> KzgAccumulator:
        """sum_j rho^j (accumulator the multi-open verifier outputs for proof j): program + one fused MSM per side"""
        if self.use_device_plan:
            return self._accumulate_new_device(instances, proofs, rho, False)[0]
        m = len(proofs)
        rows, lookup, _ = self.read_proofs(instances, proofs)
        prog = self.compiled.msm.program
        out = self.loader.fr_program_eval(prog, rows.tobytes(), m)
        out = np.frombuffer(out, dtype=np.uint8).reshape(m, len(prog.outputs), 32)
        nl = len(self._slots["lhs"])
        rho_b = (rho % R_MODULUS).to_bytes(32, "little")
        sides = []
        for side, sc in (("lhs", out[:, :nl]), ("rhs", out[:, nl:])):
            pts = self._points(lookup, m, side)
            off = np.arange(m + 1, dtype=np.uint64) * pts.shape[1]
            sides.append(self.loader.msm_batch_rlc(np.ascontiguousarray(sc).reshape(-1), pts.reshape(-1), off, rho_b, flags=CHECK_INPUTS))
        return KzgAccumulator(sides[0], sides[1])

    def accumulate(self, instances, proofs, rho: int) -> KzgAccumulator:
        """the new accumulators as above, then the old accumulators (proof.rs:148-158) with the following powers of rho"""
        m = len(proofs)
        acc = self.accumulate_new(instances, proofs, rho)
        old = self.old_accumulators(instances)
        if old:
            r = pow(rho, m, R_MODULUS)
            coeffs = b"".join(c.to_bytes(32, "little") for c in [1] + [r * pow(rho, i, R_MODULUS) % R_MODULUS for i in range(len(old))])
            acc = KzgAccumulator(self.loader.msm(coeffs, acc.lhs + b"".join(a.lhs for a in old), len(old) + 1),
                                 self.loader.msm(coeffs, acc.rhs + b"".join(a.rhs for a in old), len(old) + 1))
        return acc

    def verify_batch(self, instances, proofs, rho: int) -> bool:
        """True iff the fused accumulator passes `decide` (all proofs valid, up to the soundness error of the random `rho`)"""
        if self.use_device_plan and not self.protocol.accumulator_indices:
            return self._accumulate_new_device(instances, proofs, rho, True)[1]
        acc = self.accumulate(instances, proofs, rho)
        ok, _ = self.kzg.decide_batch(acc.lhs, acc.rhs, 1)
        return ok == b"\x01"

    # -- one proof as the reference's types ----------------------------------------------------------------------------------
    def read_proof(self, instances, proof: bytes) -> PlonkProof:
        """PlonkProof::read for one proof, as host values (points validated with `from_xy`)"""
        tl = self.tl
        rows, lookup, ch = self.read_proofs([instances], [proof])
        lay = self.compiled.layout
        val = lambda a: int.from_bytes(a.tobytes(), "little")
        chal = [val(ch[0, i]) for i in range(ch.shape[1])]

        def pt(o):
            b = lookup(o)[0].tobytes()
            return _from_xy(int.from_bytes(b[:32], "little"), int.from_bytes(b[32:], "little"))
        evals = [val(rows[0, lay["evaluations"] + i]) for i in range(len(self.protocol.evaluations))]
        return PlonkProof([pt(o) for o in tl.witnesses], chal[:tl.n_challenges], [pt(o) for o in tl.quotients], chal[tl.n_challenges],
                          evals, chal[tl.n_challenges + 1:], [pt(o) for o in tl.ws], self.old_accumulators([instances]))


class PlonkVerifier:
    """verifier/plonk.rs:96-134: `verify(vk, protocol, instances, proof)` = succinct verification + `decide_all`; raises
    `AssertionFailure("e(lhs, g2)·e(rhs, -s_g2) == O")` like the reference returns it."""

    def __init__(self, loader, kzg, svk_g: bytes, protocol: PlonkProtocol, scheme: str = "gwc19", limbs: Tuple[int, int] = (4, 68),
                 transcript: str = "evm"):
        self.batch = PlonkBatchVerifier(loader, kzg, svk_g, protocol, scheme, limbs, transcript)

    def read_proof(self, instances, proof: bytes) -> PlonkProof:
        return self.batch.read_proof(instances, proof)

    def succinct_verify(self, instances, proof: bytes) -> List[KzgAccumulator]:
        """PlonkSuccinctVerifier::verify (plonk.rs:57-93): [the proof's accumulator] + the old accumulators read from the instances"""
        b = self.batch
        return [b.accumulate_new([instances], [proof], 1)] + b.old_accumulators([instances])

    def verify(self, instances, proof: bytes) -> None:
        self.batch.kzg.decide_all(self.succinct_verify(instances, proof))
