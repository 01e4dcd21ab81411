"""`snark_verifier_b200.plonk` — PlonkProtocol-driven verification (verifier/plonk.rs:57-134, verifier/plonk/proof.rs:52-349) — on genuine
proofs from the from-scratch prover tests/plonk_toy.py: accept honest, reject tampered; GWC19 and SHPLONK; the three linearization
strategies; old accumulators from the instances; batches fused by a random linear combination.

CPU: the batch verifier runs over a stand-in loader whose four device entry points are the CPU oracle (program interpreter, Keccak
model, naive MSM, oracle pairing), which checks the compiled program and the transcript schedule without a GPU.  GPU (`-m gpu`):
the same proofs through the CUDA loader; device scalars and accumulators must equal the CPU stand-in's byte for byte."""
import random

import numpy as np
import pytest

import oracle
import snark_verifier_b200 as sv
from oracle import bn254_model as m
from oracle import evm_transcript as et
from oracle import poseidon_model as pos
from oracle.plonk_eval_model import run_program
from snark_verifier_b200 import pcs, plonk
from snark_verifier_b200.plonk import MINUS_VANISHING_TIMES_QUOTIENT, WITHOUT_CONSTANT

import plonk_toy as T

R = m.R
le = m.fe_to_le
VARIANTS = [None, WITHOUT_CONSTANT, MINUS_VANISHING_TIMES_QUOTIENT]
SCHEMES = ["gwc19", "bdfg21"]


class OracleLoader:
    """CPU stand-in for the CudaLoader methods plonk.PlonkBatchVerifier calls (TEST ONLY)."""
    fmt = sv.CANONICAL

    def evm_transcript_challenges(self, streams, stream_len, seg_end, mm):
        out = b""
        for j in range(mm):
            out += b"".join(le(c) for c in et.challenges_for_stream(streams[j * stream_len:(j + 1) * stream_len], seg_end))
        return out

    def poseidon_transcript_challenges(self, elements, stream_len, seg_end, mm):
        out = b""
        for j in range(mm):
            el = [int.from_bytes(elements[32 * (j * stream_len + i):32 * (j * stream_len + i + 1)], "little") for i in range(stream_len)]
            out += b"".join(le(c) for c in pos.challenges_for_elements(el, seg_end))
        return out

    def g1_decompress(self, compressed, n, want_elements=True):
        pts, els, valid = b"", b"", b""
        for i in range(n):
            v = int.from_bytes(compressed[32 * i:32 * i + 32], "little")
            sign, ident, x = v >> 255, (v >> 254) & 1, v & ((1 << 254) - 1)
            y = pow((x * x * x + 3) % m.P, (m.P + 1) // 4, m.P)
            ok = not ident and x < m.P and y * y % m.P == (x * x * x + 3) % m.P
            if ok and (y & 1) != sign:
                y = (-y) % m.P
            if not ok:
                pts, els, valid = pts + bytes(64), els + bytes(64), valid + b"\x00"
            else:
                pts, els, valid = pts + le(x) + le(y), els + le(x % R) + le(y % R), valid + b"\x01"
        return pts, els, valid

    def fr_program_eval(self, program, inputs, mm):
        out = b""
        ni = program.n_inputs
        for j in range(mm):
            row = [int.from_bytes(inputs[32 * (j * ni + i):32 * (j * ni + i + 1)], "little") for i in range(ni)]
            out += b"".join(le(v) for v in run_program(program.instrs, program.n_regs, program.consts, row, program.outputs))
        return out

    def msm(self, scalars, points, n, flags=0):
        return oracle.msm_native(bytes(scalars), bytes(points), n)

    def msm_batch_rlc(self, scalars, points, offsets, rho, flags=0):
        s, p = bytes(np.asarray(scalars).tobytes()), bytes(np.asarray(points).tobytes())
        rho_i, acc, coeff = int.from_bytes(rho, "little"), bytes(64), 1
        for j in range(len(offsets) - 1):
            lo, hi = int(offsets[j]), int(offsets[j + 1])
            for i in range(lo, hi):
                x, y = int.from_bytes(p[64 * i:64 * i + 32], "little"), int.from_bytes(p[64 * i + 32:64 * i + 64], "little")
                if x >= m.P or y >= m.P or ((x, y) != (0, 0) and (y * y - x * x * x - 3) % m.P):
                    raise sv.Error("point is not a valid G1Affine")          # what CHECK_INPUTS reports on the device
            sc = b"".join(le(int.from_bytes(s[32 * i:32 * i + 32], "little") * coeff % R) for i in range(lo, hi))
            acc = oracle.g1_add(acc, oracle.msm_native(sc, p[64 * lo:64 * hi], hi - lo))
            coeff = coeff * rho_i % R
        return acc


class OracleKzg:
    def __init__(self, srs):
        self.srs = srs

    def decide_batch(self, lhs, rhs, n, want_gt=False):
        acc, _ = oracle.kzg_decide_batch(bytes(lhs), bytes(rhs), n, self.srs.g2, self.srs.s_g2)
        return acc, None

    def decide_all(self, accs):
        for a in accs:
            if self.decide_batch(a.lhs, a.rhs, 1)[0] != b"\x01":
                raise sv.AssertionFailure(sv.KzgAs.ASSERTION)


@pytest.fixture(scope="module")
def srs():
    return T.Srs(3)


@pytest.fixture(scope="module")
def circuit():
    return T.Circuit(4, 11, [5, 7])


def cpu_verifier(srs, protocol, scheme, transcript="evm"):
    return plonk.PlonkVerifier(OracleLoader(), OracleKzg(srs), T.GEN, protocol, scheme, transcript=transcript)


# ---- CPU ------------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("scheme", SCHEMES)
@pytest.mark.parametrize("variant", VARIANTS, ids=["no_linearization", "without_constant", "minus_vanishing_times_quotient"])
def test_honest_proof_accepts_and_every_tampering_rejects(srs, circuit, variant, scheme):
    protocol = T.make_protocol(circuit, srs, variant)
    v = cpu_verifier(srs, protocol, scheme)
    inst = [circuit.public]
    v.verify(inst, T.prove(circuit, protocol, srs, scheme))
    for tamper in ("evaluation", "witness", "opening"):
        with pytest.raises(sv.AssertionFailure, match="e\\(lhs, g2\\)"):
            v.verify(inst, T.prove(circuit, protocol, srs, scheme, tamper=tamper))
    with pytest.raises(sv.AssertionFailure):                   # a different public input
        v.verify([[5, 8]], T.prove(circuit, protocol, srs, scheme))
    other = T.make_protocol(circuit, srs, variant, initial_state=0xBEEF)
    with pytest.raises(sv.AssertionFailure):                   # another transcript initial state (proof.rs:65-67)
        cpu_verifier(srs, other, scheme).verify(inst, T.prove(circuit, protocol, srs, scheme))


@pytest.mark.parametrize("scheme", SCHEMES)
@pytest.mark.parametrize("variant", [None, WITHOUT_CONSTANT], ids=["no_linearization", "without_constant"])
def test_poseidon_transcript_proofs_accept_and_reject(srs, circuit, variant, scheme):
    """The same protocol over the native PoseidonTranscript (transcript/halo2.rs:201-274): compressed points, little-endian scalars."""
    protocol = T.make_protocol(circuit, srs, variant)
    v = cpu_verifier(srs, protocol, scheme, "poseidon")
    inst = [circuit.public]
    proof = T.prove(circuit, protocol, srs, scheme, transcript="poseidon")
    assert len(proof) == 32 * len(v.batch.tl.items) < len(T.prove(circuit, protocol, srs, scheme))
    v.verify(inst, proof)
    for tamper in ("evaluation", "witness", "opening"):
        with pytest.raises(sv.AssertionFailure):
            v.verify(inst, T.prove(circuit, protocol, srs, scheme, tamper=tamper, transcript="poseidon"))
    with pytest.raises(sv.AssertionFailure):                   # an EVM-transcript proof's challenges do not verify under Poseidon ...
        cpu_verifier(srs, protocol, scheme, "evm").verify(inst, T.prove(circuit, protocol, srs, scheme, tamper="evaluation"))
    bad = bytearray(proof)
    bad[31] |= 0x40                                            # identity flag on a witness commitment (halo2.rs:226-241 rejects it)
    with pytest.raises(plonk.TranscriptError, match="curve point"):
        v.verify(inst, bytes(bad))
    bad = bytearray(proof)
    bad[0] ^= 1                                                # x with no square root of x^3 + 3 (or another point): reject or fail the pairing
    with pytest.raises((plonk.TranscriptError, sv.AssertionFailure)):
        v.verify(inst, bytes(bad))
    p = v.read_proof(inst, proof)
    q = cpu_verifier(srs, protocol, scheme, "evm").read_proof(inst, T.prove(circuit, protocol, srs, scheme))
    assert p.witnesses == q.witnesses and p.challenges != q.challenges          # same commitments, other Fiat-Shamir


def test_read_proof_matches_the_prover_transcript_and_error_behaviour(srs, circuit):
    protocol = T.make_protocol(circuit, srs)
    v = cpu_verifier(srs, protocol, "gwc19")
    proof = T.prove(circuit, protocol, srs, "gwc19")
    p = v.read_proof([circuit.public], proof)
    assert len(p.witnesses) == 3 and len(p.quotients) == 2 and len(p.challenges) == 1 and len(p.pcs_challenges) == 2
    assert len(p.evaluations) == len(protocol.evaluations) and len(p.pcs_points) == 2          # two distinct rotations
    tl = v.batch.tl
    stream = (protocol.transcript_initial_state).to_bytes(32, "big") + b"".join(x.to_bytes(32, "big") for x in circuit.public) + proof
    chal = et.challenges_for_stream(stream, [32 * e for e in tl.seg_end])
    assert [*p.challenges, p.z, *p.pcs_challenges] == chal
    with pytest.raises(plonk.InvalidInstances):                # proof.rs:69-77
        v.verify([[1, 2, 3]], proof)
    with pytest.raises(plonk.TranscriptError):                 # truncated proof
        v.verify([circuit.public], proof[:-32])
    bad = bytearray(proof)
    bad[32 * (tl.evaluations - tl.proof_start):32 * (tl.evaluations - tl.proof_start) + 32] = R.to_bytes(32, "big")
    with pytest.raises(plonk.TranscriptError, match="scalar"):  # evaluation >= r (transcript/evm.rs:230-245)
        v.verify([circuit.public], bytes(bad))
    bad = bytearray(proof)
    bad[63] ^= 1                                               # witness point off the curve
    with pytest.raises(sv.Error):
        v.verify([circuit.public], bytes(bad))
    with pytest.raises(plonk.TranscriptError, match="curve"):
        v.read_proof([circuit.public], bytes(bad))


def test_protocol_langranges_and_layout(srs, circuit):
    protocol = T.make_protocol(circuit, srs)
    assert protocol.langranges() == [0, 1]                     # two instance rows, rotation 0 only (protocol.rs:77-106)
    tl = plonk.TranscriptLayout(protocol, "bdfg21")
    assert tl.proof_start == 3 and tl.witnesses == [3, 5, 7] and tl.quotients == [9, 11]
    assert tl.seg_end[:2] == [9, 13] and tl.ws == [tl.evaluations + 10, tl.evaluations + 12]
    compiled = plonk.compile_plonk_verifier(protocol, "gwc19")
    slots = compiled.msm.lhs_slots
    assert slots[0] == ("g",) and [s for s in slots if s[0] == "pre"] == [("pre", i) for i in range(6)]
    assert [s for s in slots if s[0] == "quot"] == [("quot", 0), ("quot", 1)] and compiled.msm.rhs_slots == [("w", 0), ("w", 1)]
    bad = T.make_protocol(circuit, srs)
    bad.evaluations = bad.evaluations[:-1]                     # an evaluation the numerator needs is missing and has rotation != 0 ...
    bad.queries = bad.queries[:-2] + bad.queries[-1:]
    bad.evaluations.pop(7)                                     # ... a(wX): no commitment can stand in for it
    with pytest.raises(plonk.InvalidProtocol):
        plonk.compile_plonk_verifier(bad, "gwc19")


def accumulator_circuit(srs, seed=5):
    """A circuit whose instance column carries an OLD accumulator as 4 x 4 limbs of 68 bits (pcs/kzg/accumulator.rs:57-81)."""
    rnd = random.Random(seed)
    a = rnd.randrange(1, R)
    old = sv.KzgAccumulator(oracle.g1_mul(T.GEN, le(a * srs.s % R)), oracle.g1_mul(T.GEN, le(a)))
    limbs = pcs.LimbsEncoding(4, 68).to_repr(old)
    circ = T.Circuit(5, seed, [9] + limbs)
    return circ, old, [[(0, 1 + i) for i in range(16)]]


@pytest.mark.parametrize("scheme", SCHEMES)
def test_old_accumulators_are_read_from_the_instances_and_decided(srs, scheme):
    circ, old, idx = accumulator_circuit(srs)
    protocol = T.make_protocol(circ, srs, accumulator_indices=idx)
    v = cpu_verifier(srs, protocol, scheme)
    proof = T.prove(circ, protocol, srs, scheme)
    accs = v.succinct_verify([circ.public], proof)
    assert len(accs) == 2 and (accs[1].lhs, accs[1].rhs) == (old.lhs, old.rhs)
    v.verify([circ.public], proof)
    # an invalid old accumulator (lhs != s rhs) makes decide_all fail although the proof itself is fine
    bad_old = sv.KzgAccumulator(oracle.g1_add(old.lhs, T.GEN), old.rhs)
    circ2 = T.Circuit(5, 5, [9] + pcs.LimbsEncoding(4, 68).to_repr(bad_old))
    protocol2 = T.make_protocol(circ2, srs, accumulator_indices=idx)
    v2 = cpu_verifier(srs, protocol2, scheme)
    proof2 = T.prove(circ2, protocol2, srs, scheme)
    with pytest.raises(sv.AssertionFailure):
        v2.verify([circ2.public], proof2)
    assert v2.batch.verify_batch([[circ2.public]], [proof2], 0x1234567) is False


def make_batch(srs, variant, scheme, count, seed=100):
    circs = [T.Circuit(4, seed + j, [j + 1, 3 * j + 2]) for j in range(count)]
    # ONE protocol: the batch shares the fixed columns, so every circuit of the batch is built over the first one's selectors
    base = circs[0]
    batch = []
    for j in range(count):
        c = T.Circuit(4, seed, [j + 1, 3 * j + 2]) if j else base   # same seed -> same selector columns, other public inputs
        batch.append(c)
    protocol = T.make_protocol(base, srs, variant)
    for c in batch:
        assert [srs.commit(f) for f in c.fixed] == protocol.preprocessed
    return protocol, batch


@pytest.mark.parametrize("scheme", SCHEMES)
def test_batch_rlc_accepts_honest_and_rejects_one_bad_proof(srs, scheme):
    protocol, batch = make_batch(srs, None, scheme, 5)
    bv = plonk.PlonkBatchVerifier(OracleLoader(), OracleKzg(srs), T.GEN, protocol, scheme)
    insts = [[c.public] for c in batch]
    proofs = [T.prove(c, protocol, srs, scheme) for c in batch]
    rho = 0xA5A5A5A5DEADBEEF1234
    assert bv.verify_batch(insts, proofs, rho) is True
    single = plonk.PlonkVerifier(OracleLoader(), OracleKzg(srs), T.GEN, protocol, scheme)
    accs = [single.succinct_verify(i, p)[0] for i, p in zip(insts, proofs)]
    fused = bv.accumulate(insts, proofs, rho)
    powers = b"".join(le(pow(rho, j, R)) for j in range(len(accs)))
    assert fused.lhs == oracle.msm_native(powers, b"".join(a.lhs for a in accs), len(accs))
    assert fused.rhs == oracle.msm_native(powers, b"".join(a.rhs for a in accs), len(accs))
    proofs[3] = T.prove(batch[3], protocol, srs, scheme, tamper="evaluation")
    assert bv.verify_batch(insts, proofs, rho) is False


# ---- GPU ------------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def gpu(srs):
    L = sv.CudaLoader(0)
    kz = sv.KzgAs(L, sv.KzgDecidingKey(T.GEN, srs.g2, srs.s_g2))
    yield L, kz
    L.close()


@pytest.mark.gpu
@pytest.mark.parametrize("scheme", SCHEMES)
@pytest.mark.parametrize("variant", VARIANTS, ids=["no_linearization", "without_constant", "minus_vanishing_times_quotient"])
def test_device_verifier_accepts_rejects_and_equals_the_cpu_stand_in(gpu, srs, circuit, variant, scheme):
    L, kz = gpu
    protocol = T.make_protocol(circuit, srs, variant)
    v = plonk.PlonkVerifier(L, kz, T.GEN, protocol, scheme)
    ref = cpu_verifier(srs, protocol, scheme)
    inst = [circuit.public]
    proof = T.prove(circuit, protocol, srs, scheme)
    v.verify(inst, proof)
    got, exp = v.succinct_verify(inst, proof)[0], ref.succinct_verify(inst, proof)[0]
    assert (got.lhs, got.rhs) == (exp.lhs, exp.rhs)
    assert v.batch.use_device_plan                              # ^ the ONE-call device pipeline (csrc/plonk_batch.cu); below: step by step
    v.batch.use_device_plan = False
    try:
        step = v.succinct_verify(inst, proof)[0]
        assert (step.lhs, step.rhs) == (exp.lhs, exp.rhs)
        rows_d, _, ch_d = v.batch.read_proofs([inst], [proof])
    finally:
        v.batch.use_device_plan = True
    rows_c, _, ch_c = ref.batch.read_proofs([inst], [proof])
    assert rows_d.tobytes() == rows_c.tobytes() and ch_d.tobytes() == ch_c.tobytes()
    for tamper in ("evaluation", "witness", "opening"):
        with pytest.raises(sv.AssertionFailure):
            v.verify(inst, T.prove(circuit, protocol, srs, scheme, tamper=tamper))
    bad = bytearray(proof)
    bad[63] ^= 1                                               # off-curve witness: the device MSM's input check reports it
    with pytest.raises(plonk.TranscriptError, match="curve point"):
        v.verify(inst, bytes(bad))
    tl = v.batch.tl
    bad = bytearray(proof)
    o = 32 * (tl.evaluations - tl.proof_start)
    bad[o:o + 32] = R.to_bytes(32, "big")                      # an evaluation >= r: read_scalar rejects it, on the device
    with pytest.raises(plonk.TranscriptError, match="scalar"):
        v.verify(inst, bytes(bad))
    v.batch.close()


@pytest.mark.gpu
@pytest.mark.parametrize("scheme", SCHEMES)
def test_device_verifier_over_the_poseidon_transcript(gpu, srs, circuit, scheme):
    """compressed points parsed and validated on the device, Poseidon challenges on the device, then the same program / MSMs / pairing"""
    L, kz = gpu
    protocol = T.make_protocol(circuit, srs)
    v = plonk.PlonkVerifier(L, kz, T.GEN, protocol, scheme, transcript="poseidon")
    ref = cpu_verifier(srs, protocol, scheme, "poseidon")
    inst = [circuit.public]
    proof = T.prove(circuit, protocol, srs, scheme, transcript="poseidon")
    v.verify(inst, proof)
    rows_d, _, ch_d = v.batch.read_proofs([inst], [proof])
    rows_c, _, ch_c = ref.batch.read_proofs([inst], [proof])
    assert rows_d.tobytes() == rows_c.tobytes() and ch_d.tobytes() == ch_c.tobytes()
    exp = ref.succinct_verify(inst, proof)[0]
    for _ in range(2):                                          # the one-call device pipeline (twice: plan state must not leak between calls)
        got = v.succinct_verify(inst, proof)[0]
        assert v.batch.use_device_plan and (got.lhs, got.rhs) == (exp.lhs, exp.rhs)
    v.batch.use_device_plan = False                             # ... and step by step through the separate entry points
    try:
        step = v.succinct_verify(inst, proof)[0]
        assert (step.lhs, step.rhs) == (exp.lhs, exp.rhs)
    finally:
        v.batch.use_device_plan = True
    for tamper in ("evaluation", "witness", "opening"):
        with pytest.raises(sv.AssertionFailure):
            v.verify(inst, T.prove(circuit, protocol, srs, scheme, tamper=tamper, transcript="poseidon"))
    bad = bytearray(proof)
    bad[31] |= 0x40
    with pytest.raises(plonk.TranscriptError, match="curve point"):
        v.verify(inst, bytes(bad))
    o = 32 * [i for i, (kind, _) in enumerate(v.batch.tl.items) if kind == "scalar"][0]
    bad = bytearray(proof)
    bad[o:o + 32] = R.to_bytes(32, "little")                   # an evaluation >= r
    with pytest.raises(plonk.TranscriptError, match="scalar"):
        v.verify(inst, bytes(bad))
    mm = 300                                                   # a fused batch with one tampered proof
    proofs = [proof] * mm
    assert v.batch.verify_batch([inst] * mm, proofs, 0xABCDEF0123456789) is True
    proofs[77] = T.prove(circuit, protocol, srs, scheme, tamper="opening", transcript="poseidon")
    assert v.batch.verify_batch([inst] * mm, proofs, 0xABCDEF0123456789) is False


@pytest.mark.gpu
@pytest.mark.parametrize("scheme", SCHEMES)
def test_device_batch_of_256_proofs_with_old_accumulators(gpu, srs, scheme):
    L, kz = gpu
    circ, old, idx = accumulator_circuit(srs)
    protocol = T.make_protocol(circ, srs, accumulator_indices=idx)
    bv = plonk.PlonkBatchVerifier(L, kz, T.GEN, protocol, scheme)
    proof = T.prove(circ, protocol, srs, scheme)
    mm = 256
    insts, proofs = [[circ.public]] * mm, [proof] * mm
    rho = 0x1F2E3D4C5B6A79880011
    assert bv.verify_batch(insts, proofs, rho) is True
    bad = list(proofs)
    bad[200] = T.prove(circ, protocol, srs, scheme, tamper="opening")
    assert bv.verify_batch(insts, bad, rho) is False
    ref = plonk.PlonkBatchVerifier(OracleLoader(), OracleKzg(srs), T.GEN, protocol, scheme)
    a, b = bv.accumulate(insts[:3], proofs[:3], rho), ref.accumulate(insts[:3], proofs[:3], rho)
    assert (a.lhs, a.rhs) == (b.lhs, b.rhs)


# ---- committed fixture (what bench.py's "real proofs" leg replicates) -----------------------------------------------------------
def fixture_verifier(golden, loader, kzg_factory, scheme, transcript="evm"):
    fx = golden("plonk_proofs")
    H = bytes.fromhex
    protocol = plonk.simple_plonk_protocol(fx["k"], [H(p) for p in fx["preprocessed"]], fx["num_public"], None, fx["initial_state"])
    bv = plonk.PlonkBatchVerifier(loader, kzg_factory(H(fx["svk_g"]), H(fx["g2"]), H(fx["s_g2"])), H(fx["svk_g"]), protocol, scheme, transcript=transcript)
    entries = [([[int(v) for v in col] for col in e["instances"]], H(e["proof"]), e["valid"]) for e in fx[scheme if transcript == "evm" else scheme + "_" + transcript]]
    return bv, entries


@pytest.mark.parametrize("scheme,transcript", [("gwc19", "evm"), ("bdfg21", "evm"), ("bdfg21", "poseidon")])
def test_golden_proofs_fixture_on_cpu(golden, scheme, transcript):
    class Key:
        def __init__(self, g2, s_g2):
            self.g2, self.s_g2 = g2, s_g2
    bv, entries = fixture_verifier(golden, OracleLoader(), lambda g, g2, s_g2: OracleKzg(Key(g2, s_g2)), scheme, transcript)
    for inst, proof, valid in entries[:2] + entries[-1:]:
        assert bv.verify_batch([inst], [proof], 1) is valid


@pytest.mark.gpu
@pytest.mark.parametrize("scheme,transcript", [("gwc19", "evm"), ("bdfg21", "evm"), ("bdfg21", "poseidon")])
def test_golden_proofs_fixture_on_device(golden, scheme, transcript):
    L = sv.CudaLoader(0)
    try:
        bv, entries = fixture_verifier(golden, L, lambda g, g2, s_g2: sv.KzgAs(L, sv.KzgDecidingKey(g, g2, s_g2)), scheme, transcript)
        for inst, proof, valid in entries:
            assert bv.verify_batch([inst], [proof], 1) is valid
        good = [e for e in entries if e[2]]
        insts, proofs = [e[0] for e in good] * 64, [e[1] for e in good] * 64           # 512 proofs, fused
        assert bv.verify_batch(insts, proofs, 0xFEEDFACE1234567) is True
        insts[100], proofs[100] = entries[-1][0], entries[-1][1]
        assert bv.verify_batch(insts, proofs, 0xFEEDFACE1234567) is False
    finally:
        L.close()
