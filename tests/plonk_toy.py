"""TEST INFRASTRUCTURE: a from-scratch prover for a small PLONK arithmetisation, so that `snark_verifier_b200.plonk` (the mirror of
`PlonkProof::{read, evaluations, commitments, queries}` + `PlonkVerifier::verify`, verifier/plonk/proof.rs:52-349, verifier/plonk.rs:57-134)
can be exercised end to end on genuine proofs — accept honest, reject tampered — for GWC19 and SHPLONK and for the three linearization
strategies of protocol.rs:465-470.  (second half: "a from-scratch StandardPlonk prover ... to obtain genuine end-to-end
fixtures for config 1".  A literal Halo2 proof cannot be produced here — no Rust toolchain; this prover writes the transcript exactly
as `PlonkProof::read` consumes it, for a protocol object built by hand the way system/halo2.rs would describe such a circuit.)

Arithmetisation (n = 2^k rows; polynomial indices as the reference orders them: preprocessed, instance, witness, quotient):
  0..5  fixed  q_l, q_r, q_o, q_m, q_c, q_next          6  instance (one column)          7..9  advice a, b, c
  gate:  q_l a + q_r b + q_o c + q_m a b + q_c + instance = 0          next:  q_next (a(wX) - c(X)) = 0
  numerator = DistributePowers([gate, next], alpha)      quotient t = numerator / (X^n - 1), two chunks of n coefficients
The SRS secret s is KNOWN to the prover (commit(f) = [f(s)] G), so commitments are scalar arithmetic; every polynomial identity,
division and opening is real.  Nothing here is zero-knowledge (no blinding rows) — irrelevant to the verifier."""
import random

import oracle
from oracle import bn254_model as m
from oracle.evm_transcript import EvmTranscript
from oracle.poseidon_model import PoseidonTranscript
from snark_verifier_b200 import pcs
from snark_verifier_b200.plonk import MINUS_VANISHING_TIMES_QUOTIENT, WITHOUT_CONSTANT, PlonkProtocol, TranscriptLayout, simple_plonk_protocol
from snark_verifier_b200.plonk_eval import Domain, Query, Rotation
from test_pcs_mirror import p_add, p_divexact, p_eval, p_interpolate, p_mul, p_scale, vanishing

R = m.R
le = m.fe_to_le
GEN = m.g1_to_bytes(m.G1_GEN)
Q_L, Q_R, Q_O, Q_M, Q_C, Q_NEXT, INSTANCE, A, B, C, QUOTIENT = range(11)


class Srs:
    def __init__(self, seed=1):
        self.s = random.Random(seed).randrange(2, R)
        self.g2 = oracle.g2_generator()
        self.s_g2 = oracle.g2_mul(self.g2, le(self.s))

    def commit(self, f):
        return oracle.g1_mul(GEN, le(p_eval(f, self.s)))


def xy(pt):
    return int.from_bytes(pt[:32], "little"), int.from_bytes(pt[32:], "little")


class Circuit:
    """A satisfied random instance of the arithmetisation.  `public`: the instance column (its row i forces a_i = -public_i)."""

    def __init__(self, k, seed, public):
        rnd = random.Random(seed)
        self.k, self.n, self.public = k, 1 << k, [v % R for v in public]
        n, npub = self.n, len(public)
        assert npub + 3 <= n
        self.domain = Domain(k)
        w = self.domain.gen
        self.xs = [pow(w, i, R) for i in range(n)]
        col = lambda: [0] * n
        ql, qr, qo, qm, qc, qn, a, b, c = (col() for _ in range(9))
        for i in range(npub):                                   # public-input rows: a_i + instance_i = 0
            ql[i] = 1
            a[i] = (-self.public[i]) % R
            b[i], c[i] = rnd.randrange(R), rnd.randrange(R)
        for i in range(npub, n - 1):
            a[i] = c[i - 1] if qn[i - 1] else rnd.randrange(R)
            b[i] = rnd.randrange(R)
            kind = rnd.randrange(3)
            if kind == 0:                                       # multiplication  a b - c = 0
                qm[i], qo[i], c[i] = 1, R - 1, a[i] * b[i] % R
            elif kind == 1:                                     # linear combination  u a + v b - c = 0
                u, v = rnd.randrange(R), rnd.randrange(R)
                ql[i], qr[i], qo[i], c[i] = u, v, R - 1, (u * a[i] + v * b[i]) % R
            else:                                               # constant  c = v
                v = rnd.randrange(R)
                qc[i], qo[i], c[i] = (-v) % R, 1, v
            qn[i] = 1 if (i + 1 < n - 1 and rnd.randrange(2)) else 0
        a[n - 1], b[n - 1], c[n - 1] = rnd.randrange(R), rnd.randrange(R), rnd.randrange(R)
        if qn[n - 2]:
            a[n - 1] = c[n - 2]
        interp = lambda ys: p_interpolate(self.xs, ys)
        self.fixed = [interp(v) for v in (ql, qr, qo, qm, qc, qn)]
        self.advice = [interp(v) for v in (a, b, c)]
        self.instance_poly = interp(self.public + [0] * (n - npub))

    def rotated(self, f, r=1):
        """f(w^r X)"""
        wr = pow(self.domain.gen, r, R)
        return [cf * pow(wr, i, R) % R for i, cf in enumerate(f)]


def make_protocol(circ: Circuit, srs: Srs, variant=None, initial_state=0xC0DE, accumulator_indices=()):
    return simple_plonk_protocol(circ.k, [srs.commit(f) for f in circ.fixed], len(circ.public), variant, initial_state, accumulator_indices)


def compress(pt: bytes) -> bytes:
    """bn256 G1 `to_bytes` (halo2curves): x little-endian, bit 255 = parity of y (a proof never carries the identity)"""
    x, y = xy(pt)
    return (x | ((y & 1) << 255)).to_bytes(32, "little")


def prove(circ: Circuit, protocol: PlonkProtocol, srs: Srs, scheme="gwc19", tamper=None, transcript="evm") -> bytes:
    """Writes the proof exactly in the order PlonkProof::read consumes it.  `tamper`: None | "evaluation" | "witness" | "opening".
    `transcript`: "evm" (Keccak EvmTranscript: 32-byte big-endian words, uncompressed points) | "poseidon" (PoseidonTranscript:
    little-endian scalars, compressed points; system/halo2/transcript/halo2.rs:276-330)."""
    n, variant = circ.n, protocol.linearization
    evm = transcript == "evm"
    tr, out = (EvmTranscript() if evm else PoseidonTranscript()), bytearray()

    def write_point(pt):
        x, y = xy(pt)
        tr.common_ec_point(x, y)
        out.extend(x.to_bytes(32, "big") + y.to_bytes(32, "big") if evm else compress(pt))

    def write_scalar(v):
        tr.common_scalar(v)
        out.extend((v % R).to_bytes(32, "big" if evm else "little"))

    if protocol.transcript_initial_state is not None:
        tr.common_scalar(protocol.transcript_initial_state)
    for v in circ.public:
        tr.common_scalar(v)
    a, b, c = circ.advice
    wit = [srs.commit(f) for f in circ.advice]
    if tamper == "witness":
        wit[1] = oracle.g1_add(wit[1], GEN)
    for pt in wit:
        write_point(pt)
    alpha = tr.squeeze_challenge()
    ql, qr, qo, qm, qc, qn = circ.fixed
    gate = p_add(p_add(p_add(p_mul(ql, a), p_mul(qr, b)), p_add(p_mul(qo, c), p_mul(p_mul(qm, a), b))), p_add(qc, circ.instance_poly))
    nxt = p_mul(qn, p_add(circ.rotated(a), p_scale(c, R - 1)))
    numer = p_add(p_scale(gate, alpha), nxt)
    t = p_divexact(numer, [R - 1] + [0] * (n - 1) + [1])
    t = t + [0] * (2 * n - len(t))
    chunks = [t[:n], t[n:2 * n]]
    for ch in chunks:
        write_point(srs.commit(ch))
    z = tr.squeeze_challenge()
    zn = pow(z, n, R)
    h = p_add(chunks[0], p_scale(chunks[1], zn))             # the polynomial behind  sum_i (z^n)^i [t_i]
    polys = {i: f for i, f in enumerate(circ.fixed)}
    polys.update({A: a, B: b, C: c})
    ev = lambda q: p_eval(polys[q.poly], circ.domain.rotate_scalar(z, q.rotation))
    az, awz, bz, cz = ev(Query(A)), ev(Query(A, Rotation(1))), ev(Query(B)), ev(Query(C))
    if variant is None:
        polys[QUOTIENT] = h
    else:
        # r(X): the selectors as polynomials, everything else evaluated — the Msm half of the numerator (proof.rs:218-250)
        g_lin = p_add(p_add(p_add(p_scale(ql, az), p_scale(qr, bz)), p_add(p_scale(qo, cz), p_scale(qm, az * bz % R))), qc)
        r_poly = p_add(p_scale(g_lin, alpha), p_scale(qn, (awz - cz) % R))
        if variant == WITHOUT_CONSTANT:
            polys[QUOTIENT], polys[QUOTIENT + 1] = h, r_poly
        else:
            polys[QUOTIENT] = p_add(r_poly, p_scale(h, (1 - zn) % R))           # numerator - quotient (z^n - 1)
    evals = [ev(q) for q in protocol.evaluations]
    if tamper == "evaluation":
        evals[1] = (evals[1] + 1) % R
    for v in evals:
        write_scalar(v)
    # ---- multi-open proof over protocol.queries ------------------------------------------------------------------------------
    queries = [pcs.Query(q.poly, circ.domain.rotate_scalar(1, q.rotation), ev(q)) for q in protocol.queries]
    if scheme == "gwc19":
        v = tr.squeeze_challenge()
        ws = []
        for st in pcs.Gwc19.query_sets(queries):                # gwc19.rs:140-160
            point = st["shift"] * z % R
            f = [0]
            for j, poly in enumerate(st["polys"]):
                f = p_add(f, p_scale(p_add(polys[poly], [(-p_eval(polys[poly], point)) % R]), pow(v, j, R)))
            ws.append(srs.commit(p_divexact(f, [(-point) % R, 1])))
        if tamper == "opening":
            ws[0] = oracle.g1_add(ws[0], GEN)
        for w in ws:
            write_point(w)
        tr.squeeze_challenge()                                   # u
    else:
        mu, gamma = tr.squeeze_challenge(), tr.squeeze_challenge()
        sets = pcs.Bdfg21.query_sets(queries)                    # the prover groups exactly like the verifier (bdfg21.rs:123-175)
        hh, terms = [0], []
        for kk, st in enumerate(sets):
            pts = [z * sh % R for sh in st["shifts"]]
            zs = vanishing(pts)
            for i, (poly, evs) in enumerate(zip(st["polys"], st["evals"])):
                r_int = p_interpolate(pts, evs)
                coeff = pow(gamma, kk, R) * pow(mu, i, R) % R
                hh = p_add(hh, p_scale(p_divexact(p_add(polys[poly], p_scale(r_int, R - 1)), zs), coeff))
                terms.append((coeff, poly, r_int, zs))
        w1 = srs.commit(hh)
        if tamper == "opening":
            w1 = oracle.g1_add(w1, GEN)
        write_point(w1)
        z_prime = tr.squeeze_challenge()
        zs1 = p_eval(terms[0][3], z_prime)
        big_l = p_scale(hh, (-zs1) % R)
        for coeff, poly, r_int, zs in terms:
            scale = coeff * zs1 % R * pow(p_eval(zs, z_prime), -1, R) % R
            big_l = p_add(big_l, p_scale(p_add(polys[poly], [(-p_eval(r_int, z_prime)) % R]), scale))
        write_point(srs.commit(p_divexact(big_l, [(-z_prime) % R, 1])))
    assert len(out) == TranscriptLayout(protocol, scheme).proof_len(transcript)
    return bytes(out)
