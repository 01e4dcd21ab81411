"""the callers that shape the hot path's inputs — `Gwc19::verify` (gwc19.rs:45-82), `Bdfg21::verify`
(bdfg21.rs:51-83, coefficients :177-371) and `LimbsEncoding::from_repr` (accumulator.rs:57-81) — mirrored in
snark_verifier_b200/pcs.py and checked end to end against honest toy provers.

The SRS secret s is KNOWN to the test, so commit(f) = [f(s)]G and a prover's quotient commitments are scalar arithmetic; the SHPLONK
prover below does the real polynomial work (interpolation, exact divisions), so a wrong verifier coefficient makes the pairing
check fail.  CPU: through an oracle-backed NativeLoader stand-in.  GPU: through CudaLoader (same accumulator bytes, decide on device).
"""
import random

import pytest

import oracle
import snark_verifier_b200 as sv
from oracle import bn254_model as m
from snark_verifier_b200 import pcs
from test_pcs_boundary import GEN, OracleNativeLoader, gwc19_verify, make_fixture

R, P = m.R, m.P
le = m.fe_to_le


# ---- tiny polynomial arithmetic over Fr (coefficients, low degree first) ----------------------------------------------------
def p_eval(f, x):
    acc = 0
    for c in reversed(f):
        acc = (acc * x + c) % R
    return acc


def p_add(f, g):
    n = max(len(f), len(g))
    return [((f[i] if i < len(f) else 0) + (g[i] if i < len(g) else 0)) % R for i in range(n)]


def p_scale(f, k):
    return [c * k % R for c in f]


def p_mul(f, g):
    out = [0] * (len(f) + len(g) - 1)
    for i, a in enumerate(f):
        for j, b in enumerate(g):
            out[i + j] = (out[i + j] + a * b) % R
    return out


def p_divexact(f, g):
    f = list(f)
    q = [0] * (len(f) - len(g) + 1)
    inv = pow(g[-1], -1, R)
    for i in range(len(q) - 1, -1, -1):
        q[i] = f[i + len(g) - 1] * inv % R
        for j, b in enumerate(g):
            f[i + j] = (f[i + j] - q[i] * b) % R
    assert not any(f), "division must be exact"
    return q


def p_interpolate(xs, ys):
    out = [0]
    for j, (xj, yj) in enumerate(zip(xs, ys)):
        num, den = [1], 1
        for i, xi in enumerate(xs):
            if i != j:
                num = p_mul(num, [(-xi) % R, 1])
                den = den * (xj - xi) % R
        out = p_add(out, p_scale(num, yj * pow(den, -1, R) % R))
    return out


def vanishing(xs):
    z = [1]
    for x in xs:
        z = p_mul(z, [(-x) % R, 1])
    return z


# ---- SHPLONK fixture (prover follows eprint 2020/081 §4 as halo2's shplonk does; verifier = Bdfg21::verify) ------------------
def make_shplonk_fixture(seed=0, k=5, tamper=False, srs_seed=None, shuffle_seed=None):
    rnd = random.Random(seed)
    s = rnd.randrange(2, R)
    if srs_seed is not None:                                 # several proofs under ONE SRS (batch verification)
        s = random.Random(srs_seed).randrange(2, R)
    g2 = oracle.g2_generator()
    s_g2 = oracle.g2_mul(g2, le(s))
    commit = lambda v: oracle.g1_mul(GEN, le(v % R))
    omega = pow(7, (R - 1) >> k, R)
    om_inv = pow(omega, -1, R)
    # polynomial -> the shifts it is opened at (first appearance order defines the sets: {1}, {1, w}, {1, w, w^-1}, {w^-1})
    shifts_of = [[1], [1], [1, omega], [1], [1, omega, om_inv], [omega, 1], [om_inv], [1], [om_inv]]
    polys = [[rnd.randrange(R) for _ in range(1 << k)] for _ in shifts_of]
    z = rnd.randrange(R)
    queries = []
    for j, shs in enumerate(shifts_of):                     # interleave like protocol.queries would: by polynomial, then shift
        for sh in shs:
            queries.append(pcs.Query(j, sh, p_eval(polys[j], z * sh % R)))
    (random.Random(shuffle_seed) if shuffle_seed is not None else rnd).shuffle(queries)   # a batch shares ONE query order
    mu, gamma, z_prime = rnd.randrange(R), rnd.randrange(R), rnd.randrange(R)
    sets = pcs.Bdfg21.query_sets(queries)                   # the prover groups exactly like the verifier
    h = [0]
    terms = []
    for kk, st in enumerate(sets):
        pts = [z * sh % R for sh in st["shifts"]]
        zs = vanishing(pts)
        for i, (poly, evals) in enumerate(zip(st["polys"], st["evals"])):
            r_poly = p_interpolate(pts, evals)
            coeff = pow(gamma, kk, R) * pow(mu, i, R) % R
            h = p_add(h, p_scale(p_divexact(p_add(polys[poly], p_scale(r_poly, R - 1)), zs), coeff))
            terms.append((coeff, poly, p_eval(r_poly, z_prime), p_eval(zs, z_prime)))
    zs1 = terms[0][3]
    L = p_scale(h, (-zs1) % R)
    for coeff, poly, r_at, zs_at in terms:
        scale = coeff * zs1 % R * pow(zs_at, -1, R) % R
        L = p_add(L, p_scale(p_add(polys[poly], [(-r_at) % R]), scale))
    q = p_divexact(L, [(-z_prime) % R, 1])
    if tamper:
        queries[3].eval = (queries[3].eval + 1) % R
    return dict(g2=g2, s_g2=s_g2, C=[commit(p_eval(f, s)) for f in polys], queries=queries, z=z,
                proof=pcs.Bdfg21Proof(mu, gamma, commit(p_eval(h, s)), z_prime, commit(p_eval(q, s))))


def bdfg21_accumulator(loader, fx):
    return pcs.Bdfg21.verify(loader, GEN, [sv.Msm.base(loader, c) for c in fx["C"]], fx["z"], fx["queries"], fx["proof"])


def gwc19_queries(fx):
    """test_pcs_boundary's GWC19 fixture as pcs::Query values (shift = point / z)"""
    zinv = pow(fx["points"][0], -1, R)
    shift = [p * zinv % R for p in fx["points"]]
    n = sum(len(s) for s in fx["sets"])
    return [pcs.Query(j, shift[j % 3], fx["evals"][j]) for j in range(n)]


# ---- CPU ----------------------------------------------------------------------------------------------------------------------
def test_gwc19_mirror_equals_handwritten_equation_and_decides():
    fx = make_fixture()
    L = OracleNativeLoader()
    acc = pcs.Gwc19.verify(L, GEN, [sv.Msm.base(L, c) for c in fx["C"]], fx["points"][0], gwc19_queries(fx),
                           pcs.Gwc19Proof(fx["v"], fx["W"], fx["u"]))
    ref = gwc19_verify(L, fx)                              # the equation written out by hand in test_pcs_boundary.py
    assert (acc.lhs, acc.rhs) == (ref.lhs, ref.rhs)
    assert oracle.kzg_decide(acc.lhs, acc.rhs, fx["g2"], fx["s_g2"], want_gt=False)[0]


def test_gwc19_query_sets_order():
    q = [pcs.Query(0, 5, 1), pcs.Query(1, 7, 2), pcs.Query(2, 5, 3), pcs.Query(0, 7, 4)]
    sets = pcs.Gwc19.query_sets(q)
    assert [(s["shift"], s["polys"], s["evals"]) for s in sets] == [(5, [0, 2], [1, 3]), (7, [1, 0], [2, 4])]


def test_bdfg21_query_sets_group_by_shift_set():
    q = [pcs.Query(0, 1, 10), pcs.Query(1, 1, 11), pcs.Query(1, 9, 12), pcs.Query(2, 9, 13), pcs.Query(2, 1, 14), pcs.Query(0, 1, 99)]
    sets = pcs.Bdfg21.query_sets(q)
    assert [(s["shifts"], s["polys"]) for s in sets] == [([1], [0]), ([1, 9], [1, 2])]
    assert sets[1]["evals"] == [[11, 12], [14, 13]]        # poly 2's evals re-ordered to the set's shift order (bdfg21.rs:150-160)
    assert sets[0]["evals"] == [[10]]                       # a repeated (poly, shift) keeps the first evaluation (bdfg21.rs:131-135)


def test_bdfg21_honest_proof_accepts_and_tampered_rejects_on_cpu():
    for seed in (0, 1):
        fx = make_shplonk_fixture(seed)
        acc = bdfg21_accumulator(OracleNativeLoader(), fx)
        assert oracle.kzg_decide(acc.lhs, acc.rhs, fx["g2"], fx["s_g2"], want_gt=False)[0]
    bad = make_shplonk_fixture(0, tamper=True)
    acc = bdfg21_accumulator(OracleNativeLoader(), bad)
    assert not oracle.kzg_decide(acc.lhs, acc.rhs, bad["g2"], bad["s_g2"], want_gt=False)[0]


def random_point(rnd):
    return oracle.g1_mul(GEN, le(rnd.randrange(1, R)))


def test_limbs_encoding_round_trip_and_rejections():
    rnd = random.Random(3)
    enc = pcs.LimbsEncoding(4, 68)
    acc = sv.KzgAccumulator(random_point(rnd), random_point(rnd))
    limbs = enc.to_repr(acc)
    assert len(limbs) == 16 and all(l < (1 << 68) for l in limbs)
    back = enc.from_repr(limbs)
    assert (back.lhs, back.rhs) == (acc.lhs, acc.rhs)
    ident = sv.KzgAccumulator(bytes(64), acc.rhs)                      # (0, 0) is the identity: accepted by from_xy
    assert enc.from_repr(enc.to_repr(ident)).lhs == bytes(64)
    off = list(limbs); off[0] ^= 1                                      # not on the curve
    with pytest.raises(pcs.InvalidAccumulator):
        enc.from_repr(off)
    big = list(limbs); big[3] = (P >> 204) + 1                          # x >= p
    with pytest.raises(pcs.InvalidAccumulator):
        enc.from_repr(big)
    wide = list(limbs); wide[7] = 1 << 60                               # limb wider than it should be: 2^(204 + 60) does not fit 32 bytes
    with pytest.raises(pcs.InvalidAccumulator):
        enc.from_repr(wide)


def test_limbs_golden_host(golden):
    """tests/golden/limbs.json (oracle/gen_golden_plonk.py): LimbsEncoding::from_repr accepts / rejects exactly the committed rows."""
    g = golden("limbs")
    enc = pcs.LimbsEncoding(g["limbs"], g["bits"])
    for row in g["rows"]:
        limbs = [int(v, 16) for v in row["limbs"]]
        if row["valid"]:
            acc = enc.from_repr(limbs)
            assert (acc.lhs.hex(), acc.rhs.hex()) == (row["lhs"], row["rhs"]), row["note"]
        else:
            with pytest.raises(pcs.InvalidAccumulator):
                enc.from_repr(limbs)


@pytest.mark.gpu
def test_limbs_golden_device(golden):
    g = golden("limbs")
    enc = pcs.LimbsEncoding(g["limbs"], g["bits"])
    L = sv.CudaLoader(0)
    try:
        rows = g["rows"]
        buf = b"".join((int(v, 16) % R).to_bytes(32, "little") for row in rows for v in row["limbs"])
        # a limb >= r cannot be passed as an Fr (the over-wide rows stay below r: 2^70 + ... < r)
        lhs, rhs, valid = enc.from_repr_batch(L, buf, len(rows))
        for a, row in enumerate(rows):
            assert valid[a] == row["valid"], row["note"]
            assert (lhs[64 * a:64 * a + 64].hex(), rhs[64 * a:64 * a + 64].hex()) == (row["lhs"], row["rhs"]), row["note"]
    finally:
        L.close()


def test_gwc19_msm_scalar_program_equals_host_mirror():
    """gwc19.rs:52-81 compiled to a straight-line program (SymbolicMsm over virtual registers) must produce exactly the
    (scalar, base) pairs — same order, same values — that the host mirror hands to multi_scalar_multiplication."""
    from oracle import plonk_eval_model as om
    from snark_verifier_b200 import plonk_eval as pe
    rnd = random.Random(4)
    omega = pe.root_of_unity(8)
    shifts = [1, omega, pow(omega, -1, R)]
    structure = [(j, shifts[(j * 7) % 3]) for j in range(17)] + [(3, shifts[1]), (5, shifts[2])]    # some polys at two points
    mp = pe.compile_gwc19_msm_scalars(structure, 17)
    z, v, u = (rnd.randrange(R) for _ in range(3))
    evals = [rnd.randrange(R) for _ in structure]
    C = [bytes([j + 1]) * 64 for j in range(17)]
    W = [bytes([100 + i]) * 64 for i in range(3)]
    G = bytes([255]) * 64
    calls = []

    class Recorder:
        fmt = sv.CANONICAL

        def multi_scalar_multiplication(self, pairs):
            calls.append([(int.from_bytes(sc, "little"), pt) for sc, pt in pairs])
            return bytes(64)
    L = Recorder()
    pcs.Gwc19.verify(L, G, [sv.Msm.base(L, c) for c in C], z, [pcs.Query(p, sh, e) for (p, sh), e in zip(structure, evals)],
                     pcs.Gwc19Proof(v, W, u))
    out = om.run_program(mp.program.instrs, mp.program.n_regs, mp.program.consts, [z, v, u] + evals, mp.program.outputs)
    pt = lambda sl: G if sl == ("g",) else (C[sl[1]] if sl[0] == "c" else W[sl[1]])
    nl = len(mp.lhs_slots)
    assert [(out[i], pt(sl)) for i, sl in enumerate(mp.lhs_slots)] == calls[0]
    assert [(out[nl + i], pt(sl)) for i, sl in enumerate(mp.rhs_slots)] == calls[1]
    assert nl == 21 and len(mp.rhs_slots) == 3


def test_bdfg21_msm_scalar_program_equals_host_mirror():
    """bdfg21.rs:51-83 + :177-371 compiled to a straight-line program: same (scalar, base) pairs as the host mirror, two shared
    inversions (the two `L::batch_invert` rounds)."""
    from oracle import plonk_eval_model as om
    from snark_verifier_b200 import plonk_eval as pe
    for seed in (0, 3):
        fx = make_shplonk_fixture(seed)
        structure = [(q.poly, q.shift) for q in fx["queries"]]
        mp = pe.compile_bdfg21_msm_scalars(structure, len(fx["C"]))
        assert mp.program.op_histogram()["inv"] == 2
        calls = []

        class Recorder:
            fmt = sv.CANONICAL

            def multi_scalar_multiplication(self, pairs):
                calls.append([(int.from_bytes(sc, "little"), pt) for sc, pt in pairs])
                return bytes(64)
        bdfg21_accumulator(Recorder(), fx)
        pr = fx["proof"]
        row = [fx["z"], pr.mu, pr.gamma, pr.z_prime] + [q.eval for q in fx["queries"]]
        out = om.run_program(mp.program.instrs, mp.program.n_regs, mp.program.consts, row, mp.program.outputs)
        pt = lambda sl: GEN if sl == ("g",) else (fx["C"][sl[1]] if sl[0] == "c" else (pr.w, pr.w_prime)[sl[1]])
        nl = len(mp.lhs_slots)
        assert [(out[i], pt(sl)) for i, sl in enumerate(mp.lhs_slots)] == calls[0]
        assert [(out[nl + i], pt(sl)) for i, sl in enumerate(mp.rhs_slots)] == calls[1]


def test_cpp_multiopen_mirrors_emit_the_python_mirror_pairs(tmp_path):
    """The C++ host mirrors (`BasicGwc19` in host/cuda_loader.hpp, `BasicBdfg21` in host/pcs.hpp) over a recording loader: every
    (scalar, base) pair they hand to multi_scalar_multiplication — lhs call, then rhs call — equals the Python mirror's."""
    import os
    import struct
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "pcs_mirror_test"
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-o", str(exe), os.path.join(root, "tests", "pcs_mirror_test.cpp")])

    class Recorder:
        fmt = sv.CANONICAL

        def __init__(self):
            self.calls = []

        def multi_scalar_multiplication(self, pairs):
            self.calls.append([(int.from_bytes(sc, "little"), bytes(pt)) for sc, pt in pairs])
            return bytes(64)

    def run_cpp(blob):
        inp = tmp_path / "in.bin"
        inp.write_bytes(blob)
        out = subprocess.run([str(exe), str(inp)], capture_output=True, text=True, check=True).stdout.split("\n")
        calls = []
        for line in out:
            if line.startswith("msm "):
                calls.append([])
            elif line:
                sc, pt = line.split()
                calls[-1].append((int(sc, 16), bytes.fromhex(pt)))
        return calls

    def common(scheme, commitments, queries, z):
        blob = struct.pack("<BII", scheme, len(commitments), len(queries)) + le(z)
        for q in queries:
            blob += struct.pack("<I", q.poly) + le(q.shift) + le(q.eval)
        return blob + b"".join(commitments) + GEN

    # GWC19
    fx = make_fixture(seed=8, k=5)
    qs = gwc19_queries(fx)
    rec = Recorder()
    pcs.Gwc19.verify(rec, GEN, [sv.Msm.base(rec, c) for c in fx["C"]], fx["points"][0], qs, pcs.Gwc19Proof(fx["v"], fx["W"], fx["u"]))
    blob = common(0, fx["C"], qs, fx["points"][0]) + le(fx["v"]) + le(fx["u"]) + struct.pack("<I", len(fx["W"])) + b"".join(fx["W"])
    assert run_cpp(blob) == rec.calls and len(rec.calls) == 2
    # SHPLONK
    for seed in (0, 4):
        sx = make_shplonk_fixture(seed)
        rec = Recorder()
        bdfg21_accumulator(rec, sx)
        pr = sx["proof"]
        blob = common(1, sx["C"], sx["queries"], sx["z"]) + le(pr.mu) + le(pr.gamma) + pr.w + le(pr.z_prime) + pr.w_prime
        assert run_cpp(blob) == rec.calls and len(rec.calls) == 2


def test_cpp_msm_scalar_programs_equal_python(tmp_path):
    """host/plonk_eval.hpp `compile_gwc19_msm_scalars` / `compile_bdfg21_msm_scalars` (SymbolicMsm over a ProgramBuilder in C++)
    emit, instruction for instruction and slot for slot, the programs the Python compiler emits."""
    import os
    import struct
    import subprocess
    from snark_verifier_b200 import plonk_eval as pe
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "plonk_compile_test"
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-o", str(exe), os.path.join(root, "tests", "plonk_compile_test.cpp")])
    omega = pe.root_of_unity(8)
    shifts = [1, omega, pow(omega, -1, R)]
    gwc_structure = [(j, shifts[(j * 7) % 3]) for j in range(17)] + [(3, shifts[1]), (5, shifts[2])]
    sx = make_shplonk_fixture(1)
    cases = [("gwc19", gwc_structure, 17, pe.compile_gwc19_msm_scalars),
             ("bdfg21", [(q.poly, q.shift) for q in sx["queries"]], len(sx["C"]), pe.compile_bdfg21_msm_scalars),
             ("bdfg21", [(j, shifts[j % 3]) for j in range(6)] + [(2, shifts[0]), (4, shifts[0]), (4, shifts[1])], 6, pe.compile_bdfg21_msm_scalars)]
    for mode, structure, npoly, compile_py in cases:
        inp = tmp_path / "q.bin"
        inp.write_bytes(struct.pack("<II", npoly, len(structure)) + b"".join(struct.pack("<I", p) + le(sh) for p, sh in structure))
        out = subprocess.run([str(exe), mode, str(inp)], capture_output=True, text=True, check=True).stdout.splitlines()
        mp = compile_py(structure, npoly)
        prog = mp.program
        hdr = out[0].split()
        assert (int(hdr[1]), int(hdr[3])) == (prog.n_regs, prog.n_inputs)
        assert [tuple(int(x) for x in l.split()[1:]) for l in out if l.startswith("i ")] == [tuple(i) for i in prog.instrs]
        assert [int(l.split()[1], 16) for l in out if l.startswith("c ")] == prog.consts
        assert [int(x) for x in [l for l in out if l.startswith("o")][0].split()[1:]] == prog.outputs
        slot = lambda sl: "g0" if sl == ("g",) else "%s%d" % (sl[0], sl[1])
        assert [l for l in out if l.startswith("l")][0].split()[1:] == [slot(sl) for sl in mp.lhs_slots]
        assert [l for l in out if l.startswith("r")][0].split()[1:] == [slot(sl) for sl in mp.rhs_slots]


# ---- GPU ----------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_bdfg21_batch_verifier_pipeline_on_device():
    """The SHPLONK twin of the test below: m honest proofs (real polynomial divisions) under one SRS, scalars by the device program
    (two shared inversions per proof), one fused MSM per side, one pairing; fused accumulator == RLC of the host mirror's."""
    m_proofs, srs = 12, 777
    fxs = [make_shplonk_fixture(seed=200 + j, srs_seed=srs, shuffle_seed=9) for j in range(m_proofs)]
    structure = [(q.poly, q.shift) for q in fxs[0]["queries"]]
    assert all([(q.poly, q.shift) for q in fx["queries"]] == structure for fx in fxs)

    def as_proof(fx):
        pr = fx["proof"]
        return dict(z=fx["z"], mu=pr.mu, gamma=pr.gamma, z_prime=pr.z_prime, evals=[q.eval for q in fx["queries"]],
                    commitments=fx["C"], ws=[pr.w, pr.w_prime])
    L = sv.CudaLoader(0)
    try:
        kz = sv.KzgAs(L, sv.KzgDecidingKey(GEN, fxs[0]["g2"], fxs[0]["s_g2"]))
        bv = pcs.Bdfg21BatchVerifier(L, kz, GEN, structure, len(fxs[0]["C"]))
        rho = 0x0F1E2D3C4B5A69788796A5B4C3D2E1F0
        fused = bv.accumulate([as_proof(fx) for fx in fxs], rho)
        N = OracleNativeLoader()
        per = [bdfg21_accumulator(N, fx) for fx in fxs]
        rs = b"".join(le(pow(rho, j, R)) for j in range(m_proofs))
        assert fused.lhs == oracle.msm_native(rs, b"".join(a.lhs for a in per), m_proofs)
        assert fused.rhs == oracle.msm_native(rs, b"".join(a.rhs for a in per), m_proofs)
        bv.verify_batch([as_proof(fx) for fx in fxs], rho)
        bad = [as_proof(fx) for fx in fxs]
        bad[5] = as_proof(make_shplonk_fixture(seed=205, srs_seed=srs, shuffle_seed=9, tamper=True))
        with pytest.raises(sv.AssertionFailure):
            bv.verify_batch(bad, rho)
    finally:
        L.close()


@pytest.mark.gpu
def test_gwc19_batch_verifier_pipeline_on_device():
    """BASELINE config 3 with REAL GWC19 structure: m proofs under one SRS -> MSM scalars by the device program -> one fused MSM per
    side (powers of rho) -> one pairing.  The fused accumulator must equal the RLC of the per-proof accumulators of the host
    mirror (NativeLoader fold), the batch must accept, and one tampered evaluation must make it reject."""
    m_proofs, srs = 24, 4242
    fxs = [make_fixture(seed=100 + j, k=5, srs_seed=srs) for j in range(m_proofs)]
    zinv = [pow(fx["points"][0], -1, R) for fx in fxs]
    shifts = [p * zinv[0] % R for p in fxs[0]["points"]]
    structure = [(j, shifts[j % 3]) for j in range(17)]

    def as_proof(fx):
        return dict(z=fx["points"][0], v=fx["v"], u=fx["u"], evals=[fx["evals"][j] for j in range(17)], commitments=fx["C"], ws=fx["W"])
    L = sv.CudaLoader(0)
    try:
        kz = sv.KzgAs(L, sv.KzgDecidingKey(GEN, fxs[0]["g2"], fxs[0]["s_g2"]))
        bv = pcs.Gwc19BatchVerifier(L, kz, GEN, structure, 17)
        rho = 0x1F2E3D4C5B6A79880123456789ABCDEF
        fused = bv.accumulate([as_proof(fx) for fx in fxs], rho)
        N = OracleNativeLoader()
        per = [pcs.Gwc19.verify(N, GEN, [sv.Msm.base(N, c) for c in fx["C"]], fx["points"][0], gwc19_queries(fx),
                                pcs.Gwc19Proof(fx["v"], fx["W"], fx["u"])) for fx in fxs]
        rs = b"".join(le(pow(rho, j, R)) for j in range(m_proofs))
        assert fused.lhs == oracle.msm_native(rs, b"".join(a.lhs for a in per), m_proofs)
        assert fused.rhs == oracle.msm_native(rs, b"".join(a.rhs for a in per), m_proofs)
        bv.verify_batch([as_proof(fx) for fx in fxs], rho)
        bad = [as_proof(fx) for fx in fxs]
        bad[7] = as_proof(make_fixture(seed=107, k=5, srs_seed=srs, tamper=True))
        with pytest.raises(sv.AssertionFailure):
            bv.verify_batch(bad, rho)
    finally:
        L.close()



@pytest.mark.gpu
def test_multiopen_verifiers_on_cuda_loader():
    L = sv.CudaLoader(0)
    try:
        fx = make_shplonk_fixture(2)
        acc = bdfg21_accumulator(L, fx)
        ref = bdfg21_accumulator(OracleNativeLoader(), fx)
        assert (acc.lhs, acc.rhs) == (ref.lhs, ref.rhs)               # same accumulator bytes as the NativeLoader fold
        kz = sv.KzgAs(L, sv.KzgDecidingKey(GEN, fx["g2"], fx["s_g2"]))
        kz.decide(acc)
        with pytest.raises(sv.AssertionFailure):
            kz.decide(bdfg21_accumulator(L, make_shplonk_fixture(2, tamper=True)))
        gx = make_fixture(5)
        g_acc = pcs.Gwc19.verify(L, GEN, [sv.Msm.base(L, c) for c in gx["C"]], gx["points"][0], gwc19_queries(gx),
                                 pcs.Gwc19Proof(gx["v"], gx["W"], gx["u"]))
        kz2 = sv.KzgAs(L, sv.KzgDecidingKey(GEN, gx["g2"], gx["s_g2"]))
        kz2.decide(g_acc)
    finally:
        L.close()


@pytest.mark.gpu
@pytest.mark.parametrize("fmt_mont", [False, True])
def test_device_from_repr_batch_matches_host(fmt_mont):
    rnd = random.Random(9)
    enc = pcs.LimbsEncoding(4, 68)
    L = sv.CudaLoader(0, fmt=sv.MONTGOMERY if fmt_mont else sv.CANONICAL)
    try:
        m_acc = 300
        base = [oracle.g1_mul(GEN, le(rnd.randrange(1, R))) for _ in range(8)]
        rows, expect = [], []
        for a in range(m_acc):
            acc = sv.KzgAccumulator(base[a % 8], base[(3 * a + 1) % 8])
            if a % 41 == 7:
                acc = sv.KzgAccumulator(bytes(64), acc.rhs)            # identity
            limbs = enc.to_repr(acc)
            kind = a % 10
            if kind == 3:
                limbs[4] ^= 2                                          # off the curve
            elif kind == 5:
                limbs[11] = (P >> 204) + 5                             # coordinate >= p
            elif kind == 8:
                limbs[15] = (1 << 70) + limbs[15]                      # does not fit 32 bytes
            rows.append(limbs)
            try:
                e = enc.from_repr(limbs)
                expect.append((e.lhs, e.rhs, 1))
            except pcs.InvalidAccumulator:
                expect.append((bytes(64), bytes(64), 0))
        conv = (lambda v: (v << 256) % R) if fmt_mont else (lambda v: v)
        buf = b"".join(conv(l).to_bytes(32, "little") for row in rows for l in row)
        lhs, rhs, valid = enc.from_repr_batch(L, buf, m_acc)
        qinv = pow(1 << 256, -1, P)

        def pt(b):
            if not fmt_mont:
                return b
            return b"".join((int.from_bytes(b[i:i + 32], "little") * qinv % P).to_bytes(32, "little") for i in (0, 32))
        got = [(pt(lhs[64 * a:64 * a + 64]), pt(rhs[64 * a:64 * a + 64]), valid[a]) for a in range(m_acc)]
        assert got == expect
        assert sum(v for _, _, v in expect) not in (0, m_acc)
        with pytest.raises(sv.Error):
            L.accumulators_from_limbs(buf, m_acc, 9, 68)               # usage error: too many limbs
    finally:
        L.close()
