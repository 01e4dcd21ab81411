"""GPU parity of the Pallas build of the MSM pipeline and the IPA decider (csrc/msm_pasta.cu) against the independent
Python big-int model oracle/pasta_model.py: field operations of both Pallas fields, small and skewed MSMs bit for bit, h_coeffs,
IpaAs::decide accept / reject, and the discrete-log checksum at 2^20 terms.  Bar: bit-exact."""
import random

import numpy as np
import pytest

import snark_verifier_b200 as sv
from oracle import pasta_model as pm
from snark_verifier_b200 import pasta

pytestmark = pytest.mark.gpu
le = pm.fe_to_le


@pytest.fixture(scope="module")
def loader():
    L = sv.CudaLoader(0)
    yield L
    L.close()


@pytest.fixture(scope="module")
def pallas(loader):
    return pasta.PallasLoader(loader)


def test_pallas_field_ops_both_fields(pallas):
    rnd = random.Random(3)
    for field, mod in ((0, pm.P), (1, pm.Q)):
        edge = [0, 1, 2, mod - 1, mod - 2, (1 << 254) % mod, (1 << 256) % mod]
        xs = edge + [rnd.randrange(mod) for _ in range(500)]
        ys = list(reversed(edge)) + [rnd.randrange(mod) for _ in range(500)]
        a, b, n = b"".join(map(le, xs)), b"".join(map(le, ys)), len(xs)
        assert pallas.field_op(field, 0, a, b, n) == b"".join(le(x * y % mod) for x, y in zip(xs, ys))
        assert pallas.field_op(field, 1, a, b, n) == b"".join(le((x + y) % mod) for x, y in zip(xs, ys))
        assert pallas.field_op(field, 2, a, b, n) == b"".join(le((x - y) % mod) for x, y in zip(xs, ys))
        nz = [x for x in xs if x]
        inv = pallas.field_op(field, 3, b"".join(map(le, nz)), b"".join(map(le, nz)), len(nz))
        assert inv == b"".join(le(pow(x, -1, mod)) for x in nz)


def test_pallas_group_law_known_answers(pallas):
    g = pasta.PALLAS_GENERATOR
    assert pallas.msm(le(1), g, 1) == g
    assert pallas.msm(le(2), g, 1) == pm.pt_to_bytes(pm.mul(pm.GEN, 2)) == pallas.msm(le(1) * 2, g * 2, 2)
    assert pallas.msm(le(pm.Q - 1), g, 1) == pm.pt_to_bytes(pm.neg(pm.GEN))                     # [q - 1] G = -G  (the group order)
    assert pallas.msm(le(pm.Q - 1) + le(1), g * 2, 2) == bytes(64)                               # ... + G = O
    assert pallas.msm(le(5), bytes(64), 1) == bytes(64)


@pytest.mark.parametrize("n", [1, 2, 3, 33, 257, 1500])
def test_pallas_msm_vs_model(pallas, n):
    rnd = random.Random(n)
    pts = [pm.mul(pm.GEN, rnd.randrange(1, pm.Q)) for _ in range(min(n, 40))]
    points = [pts[rnd.randrange(len(pts))] for _ in range(n)]
    scalars = [rnd.randrange(pm.Q) for _ in range(n)]
    scalars[0] = pm.Q - 1                                                                       # a 255-bit scalar: top window + carry
    if n > 2:
        scalars[1], points[2] = 0, None
    got = pallas.msm(b"".join(map(le, scalars)), b"".join(map(pm.pt_to_bytes, points)), n)
    assert got == pm.pt_to_bytes(pm.msm_naive(scalars, points))


@pytest.mark.parametrize("c", [3, 7, 13, 16])
def test_pallas_msm_every_window_size_and_skew(loader, pallas, c):
    rnd = random.Random(c)
    n = 600
    points = [pm.mul(pm.GEN, rnd.randrange(1, 1 << 40)) for _ in range(n)]
    pb = b"".join(map(pm.pt_to_bytes, points))
    loader.set_window_bits(c)
    try:
        for scalars in ([rnd.randrange(pm.Q) for _ in range(n)], [pm.Q - 1] * n, [1] * n, [(1 << 254) + 5] * n):
            assert pallas.msm(b"".join(map(le, scalars)), pb, n, flags=sv.CHECK_INPUTS) == pm.pt_to_bytes(pm.msm_naive(scalars, points))
    finally:
        loader.set_window_bits(0)


def test_pallas_rejects_invalid_inputs(pallas):
    g = pasta.PALLAS_GENERATOR
    with pytest.raises(sv.Error):                                                               # scalar >= q
        pallas.msm(le(pm.Q), g, 1, flags=sv.CHECK_INPUTS)
    with pytest.raises(sv.Error):                                                               # (1, 2) is BN254's generator, not on Pallas
        pallas.msm(le(1), le(1) + le(2), 1, flags=sv.CHECK_INPUTS)


def test_h_coeffs_and_ipa_decide(loader, pallas):
    rnd = random.Random(8)
    k = 6
    xi = [rnd.randrange(1, pm.Q) for _ in range(k)]
    h = pm.h_coeffs(xi, 1)
    assert pallas.h_coeffs(b"".join(map(le, xi)), k) == b"".join(map(le, h))
    s = rnd.randrange(pm.Q)
    assert pallas.h_coeffs(b"".join(map(le, xi)), k, le(s)) == b"".join(map(le, pm.h_coeffs(xi, s)))
    g = [pm.mul(pm.GEN, rnd.randrange(1, pm.Q)) for _ in range(1 << k)]
    u = pm.msm_naive(h, g)
    assert pm.ipa_decide(g, u, xi)
    ipa = pasta.IpaAs(loader, pasta.IpaDecidingKey(b"".join(map(pm.pt_to_bytes, g))))
    good = pasta.IpaAccumulator([le(x) for x in xi], pm.pt_to_bytes(u))
    bad_u = pasta.IpaAccumulator([le(x) for x in xi], pm.pt_to_bytes(pm.add(u, pm.GEN)))
    bad_xi = pasta.IpaAccumulator([le(x) for x in xi[:-1]] + [le((xi[-1] + 1) % pm.Q)], pm.pt_to_bytes(u))
    assert ipa.decide_batch([good, bad_u, good, bad_xi]) == b"\x01\x00\x01\x00"
    ipa.decide(good)
    ipa.decide_all([good, good])
    ipa.decide_all([])
    with pytest.raises(sv.AssertionFailure, match="U == commit"):
        ipa.decide(bad_u)
    with pytest.raises(sv.AssertionFailure):
        ipa.decide_all([good, bad_xi])
    with pytest.raises(sv.Error):                                                               # a committing key with an off-curve point
        pasta.IpaAs(loader, pasta.IpaDecidingKey(le(1) + le(2) + pm.pt_to_bytes(g[1])))


def test_synth_generators_match_model(pallas):
    import torch
    n = 300
    ds = torch.empty(n * 32, dtype=torch.uint8, device="cuda")
    dp = torch.empty(n * 64, dtype=torch.uint8, device="cuda")
    pallas.synth_scalars_device(9, 100, n, ds.data_ptr())
    pallas.synth_points_device(9, 100, n, dp.data_ptr())
    torch.cuda.synchronize()
    assert bytes(ds.cpu().numpy()) == b"".join(le(pm.synth_scalar(9, 100 + i)) for i in range(n))
    assert bytes(dp.cpu().numpy()) == b"".join(pm.pt_to_bytes(pm.mul(pm.GEN, pm.synth_point_scalar(9, 100 + i))) for i in range(n))


@pytest.mark.parametrize("logn,mode", [(16, 1), (20, 0), (20, 2)])
def test_pallas_full_size_dlog_checksum(loader, pallas, logn, mode):
    """P_i = [t_i] G  =>  MSM = [sum s_i t_i] G: size-independent check of the whole pipeline (sort, both accumulation kernels)."""
    import torch
    n = 1 << logn
    ds = torch.empty(n * 32, dtype=torch.uint8, device="cuda")
    dp = torch.empty(n * 64, dtype=torch.uint8, device="cuda")
    out = torch.zeros(64, dtype=torch.uint8, device="cuda")
    pallas.synth_scalars_device(78, 0, n, ds.data_ptr())
    pallas.synth_points_device(78, 0, n, dp.data_ptr())
    loader.set_accumulate_mode(mode)
    try:
        pallas.msm_device(ds.data_ptr(), dp.data_ptr(), n, d_out_affine=out.data_ptr())
        torch.cuda.synchronize()
    finally:
        loader.set_accumulate_mode(0)
    s = np.frombuffer(ds.cpu().numpy().tobytes(), dtype="<u8").reshape(n, 4)
    total = 0
    for i in range(n):
        si = int(s[i, 0]) | (int(s[i, 1]) << 64) | (int(s[i, 2]) << 128) | (int(s[i, 3]) << 192)
        total = (total + si * pm.synth_point_scalar(78, i)) % pm.Q
    assert bytes(out.cpu().numpy()) == pm.pt_to_bytes(pm.mul(pm.GEN, total))
