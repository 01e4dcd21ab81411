"""BASELINE config 1 at the PCS boundary: a StandardPlonk-shaped GWC19 multi-open verification — 21-term lhs MSM, 3-term rhs MSM,
one KZG decide (8 preprocessed + 6 witness + 3 quotient commitments, G, 3 opening proofs W).

A literal Halo2 proof cannot be produced here (no Rust, no halo2_proofs), so the fixture is built at the boundary the
hot path starts from: a toy SRS with a KNOWN secret s (so commit(f) = [f(s)]G), 17 random polynomials of degree < 2^8 opened at
three rotations of z, the GWC19 verifier equation of pcs/kzg/multiopen/gwc19.rs:45-82 assembled with the `Msm` algebra
(util/msm.rs) and handed to the loader's MSM + `KzgAs::decide` (decider.rs:70-82).  Shapes as in the reference's tests:
accept the valid proof, reject a tampered evaluation (system/halo2/test/kzg/native.rs:57-68, test/kzg/evm.rs:58-62).

Runs once on the CPU through an oracle-backed NativeLoader stand-in (plumbing, no GPU) and once on the B200 through CudaLoader."""
import random

import pytest

import oracle
import snark_verifier_b200 as sv
from oracle import bn254_model as m

R = m.R
le = m.fe_to_le
GEN = m.g1_to_bytes(m.G1_GEN)


class OracleNativeLoader:
    """NativeLoader stand-in: `multi_scalar_multiplication` is the literal fold of loader/native.rs:61-71 (CPU oracle)."""
    fmt = sv.CANONICAL

    def multi_scalar_multiplication(self, pairs):
        pairs = list(pairs)
        return oracle.msm_native(b"".join(s for s, _ in pairs), b"".join(p for _, p in pairs), len(pairs))


def poly_eval(coeffs, x):
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * x + c) % R
    return acc


def make_fixture(seed=0, num_polys=17, k=8, tamper=False, srs_seed=None):
    rnd = random.Random(seed)
    s = rnd.randrange(2, R)
    if srs_seed is not None:                                       # several proofs under ONE SRS (batch verification)
        s = random.Random(srs_seed).randrange(2, R)
    g2 = oracle.g2_generator()
    s_g2 = oracle.g2_mul(g2, le(s))
    polys = [[rnd.randrange(R) for _ in range(1 << k)] for _ in range(num_polys)]
    commit = lambda fs: oracle.g1_mul(GEN, le(fs % R))            # [f(s)]G  ==  sum f_i [s^i]G
    C = [commit(poly_eval(f, s)) for f in polys]
    z = rnd.randrange(R)
    omega = pow(5, (R - 1) >> k, R)                                # a 2^k-th root of unity
    points = [z, z * omega % R, z * pow(omega, -1, R) % R]        # rotations 0, +1, -1 ("last" modelled as -1)
    sets = [[j for j in range(num_polys) if j % 3 == r] for r in range(3)]
    evals = {j: poly_eval(polys[j], points[r]) for r in range(3) for j in sets[r]}
    v, u = rnd.randrange(R), rnd.randrange(R)
    W = []
    for r in range(3):
        num = sum(pow(v, i, R) * (poly_eval(polys[j], s) - evals[j]) for i, j in enumerate(sets[r])) % R
        W.append(commit(num * pow(s - points[r], -1, R)))
    if tamper:
        evals[sets[1][0]] = (evals[sets[1][0]] + 1) % R
    return dict(g2=g2, s_g2=s_g2, C=C, W=W, sets=sets, points=points, evals=evals, v=v, u=u)


def gwc19_verify(loader, fx):
    """-> KzgAccumulator(lhs, rhs) exactly as Gwc19::verify builds it (gwc19.rs:52-81)."""
    lhs = sv.Msm(loader)
    rhs = sv.Msm(loader)
    for r in range(3):
        ur = pow(fx["u"], r, R)
        eval_sum = 0
        for i, j in enumerate(fx["sets"][r]):
            vi = pow(fx["v"], i, R)
            lhs = lhs + sv.Msm.base(loader, fx["C"][j]) * (ur * vi % R)
            eval_sum = (eval_sum + vi * fx["evals"][j]) % R
        lhs = lhs + sv.Msm.constant_(loader, (-ur * eval_sum) % R)
        lhs = lhs + sv.Msm.base(loader, fx["W"][r]) * (ur * fx["points"][r] % R)
        rhs = rhs + sv.Msm.base(loader, fx["W"][r]) * ur
    assert len(lhs.bases) == 20 and lhs.constant is not None and len(rhs.bases) == 3      # 21-term / 3-term MSMs
    return sv.KzgAccumulator(lhs.evaluate(GEN), rhs.evaluate(None))


def test_config1_native_loader_plumbing_on_cpu():
    fx = make_fixture()
    acc = gwc19_verify(OracleNativeLoader(), fx)
    assert oracle.kzg_decide(acc.lhs, acc.rhs, fx["g2"], fx["s_g2"], want_gt=False)[0]
    bad = gwc19_verify(OracleNativeLoader(), make_fixture(tamper=True))
    assert not oracle.kzg_decide(bad.lhs, bad.rhs, fx["g2"], fx["s_g2"], want_gt=False)[0]


@pytest.mark.gpu
def test_config1_cuda_loader_drop_in():
    fx = make_fixture()
    L = sv.CudaLoader(0)
    try:
        acc = gwc19_verify(L, fx)
        ref = gwc19_verify(OracleNativeLoader(), fx)
        assert (acc.lhs, acc.rhs) == (ref.lhs, ref.rhs)                     # same accumulator bytes as the NativeLoader fold
        kz = sv.KzgAs(L, sv.KzgDecidingKey(GEN, fx["g2"], fx["s_g2"]))
        kz.decide(acc)
        kz.decide_all([acc, acc])
        with pytest.raises(sv.AssertionFailure):
            kz.decide(gwc19_verify(L, make_fixture(tamper=True)))
    finally:
        L.close()
