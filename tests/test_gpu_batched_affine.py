"""GPU parity of the batched-affine bucket accumulation kernels — the tree kernel (csrc/bucket_affine.cuh, accumulate mode 2) and
the chained kernel (csrc/bucket_chain.cuh, mode 4) forced at every size; modes 3 / 5 additionally run the XYZZ kernel and compare
every task result on the device (the library reports the first differing task).  Same oracle / checksum bar as tests/test_gpu_parity.py: bit-exact."""
import numpy as np
import pytest

import oracle
import snark_verifier_b200 as sv
from oracle import bn254_model as m

pytestmark = pytest.mark.gpu
le = m.fe_to_le


@pytest.fixture(scope="module")
def loader():
    L = sv.CudaLoader(0)
    yield L
    L.close()


@pytest.fixture(params=[2, 3, 4, 5], ids=["affine", "affine_selfcheck", "chain", "chain_selfcheck"])
def mode(request, loader):
    loader.set_accumulate_mode(request.param)
    yield request.param
    loader.set_accumulate_mode(0)


@pytest.mark.parametrize("n", [1, 2, 5, 33, 100, 1000, 4097, 1 << 14])
def test_msm_vs_oracle_seeded(loader, mode, n):
    s = oracle.synth_scalars(31, 0, n)
    p = oracle.synth_points(31, 0, n, 8)
    assert loader.msm(s, p, n) == oracle.msm_pippenger(s, p, n, 8)


@pytest.mark.parametrize("c", [2, 3, 4, 6, 8, 13])
@pytest.mark.parametrize("glv", [1, 2], ids=["glv", "noglv"])
def test_long_lists_small_windows(loader, mode, c, glv):
    """Small windows make every bucket list long (n / 2^(c-1) points): many tree levels, several batches per level, split tasks."""
    n = 6000
    s = oracle.synth_scalars(32, 0, n)
    p = oracle.synth_points(32, 0, n, 8)
    loader.set_window_bits(c)
    loader.set_glv_mode(glv)
    try:
        assert loader.msm(s, p, n) == oracle.msm_pippenger(s, p, n, 8)
    finally:
        loader.set_window_bits(0)
        loader.set_glv_mode(0)


def test_exceptional_pairs_equal_opposite_identity(loader, mode):
    """Neighbours in a bucket list that are equal (tangent case), opposite (sum is the identity) or the identity itself."""
    n = 4096
    g = m.g1_mul(m.G1_GEN, 987654321)
    gb = m.g1_to_bytes(g)
    ngb = m.g1_to_bytes(m.g1_neg(g))
    ident = bytes(64)
    k = 0x1F3
    # (a) the same point n times with the same scalar: every pair at every level is a doubling
    assert loader.msm(le(k) * n, gb * n, n) == oracle.g1_mul(gb, le(k * n % m.R))
    # (b) P and -P alternate: every level-0 pair cancels
    assert loader.msm(le(k) * n, (gb + ngb) * (n // 2), n) == ident
    # (c) same, but one extra P survives; identities sprinkled in
    pts = (gb + ngb) * (n // 2 - 1) + ident + gb
    assert loader.msm(le(k) * n, pts, n) == oracle.g1_mul(gb, le(k))
    # (d) P with scalar k and P with scalar r - k (a negative digit pattern of the same point)
    sc = (le(k) + le(m.R - k)) * (n // 2)
    assert loader.msm(sc, gb * n, n) == ident
    # (e) a mix against the oracle: few distinct points, random scalars
    pool = [m.g1_to_bytes(m.g1_mul(m.G1_GEN, 1000 + i)) for i in range(5)] + [ident]
    rng = np.random.default_rng(5)
    pts = b"".join(pool[i] for i in rng.integers(0, len(pool), n))
    s = oracle.synth_scalars(33, 0, n)
    assert loader.msm(s, pts, n) == oracle.msm_pippenger(s, pts, n, 8)


def test_skewed_scalars_single_bucket(loader, mode):
    n = 5000
    p = oracle.synth_points(34, 0, n, 8)
    for val in (1, 2, 0xFFFF, m.R - 1):
        s = le(val) * n
        assert loader.msm(s, p, n) == oracle.msm_pippenger(s, p, n, 8), val


def test_pair_msm_accumulate_256(loader, mode, golden):
    """KzgAs::verify (accumulation.rs:41-63): two MSMs sharing the scalars — the kernel's base-set dimension."""
    g = golden("pairing")
    H = bytes.fromhex
    kz = sv.KzgAs(loader, sv.KzgDecidingKey(m.g1_to_bytes(m.G1_GEN), H(g["g2_generator"]), H(g["s_g2"])))
    n = 256
    gen = m.g1_to_bytes(m.G1_GEN)
    lhs = b"".join(oracle.g1_mul(gen, le(1000 + 7 * i)) for i in range(n))
    rhs = b"".join(oracle.g1_mul(gen, le(5000 + 3 * i)) for i in range(n))
    r = le(0x1234567890ABCDEF1234567890ABCDEF % m.R)
    got = kz.verify([sv.KzgAccumulator(lhs[64 * i:64 * i + 64], rhs[64 * i:64 * i + 64]) for i in range(n)], r)
    el, er = oracle.kzg_accumulate(lhs, rhs, n, r)
    assert (got.lhs, got.rhs) == (el, er)


@pytest.mark.parametrize("logn", [18, 20])
def test_full_size_dlog_checksum(loader, mode, logn):
    import torch
    n = 1 << logn
    ds = torch.empty(n * 32, dtype=torch.uint8, device="cuda")
    dp = torch.empty(n * 64, dtype=torch.uint8, device="cuda")
    out = torch.zeros(64, dtype=torch.uint8, device="cuda")
    loader.synth_scalars_device(78, 0, n, ds.data_ptr())
    loader.synth_points_device(78, 0, n, dp.data_ptr())
    loader.msm_device(ds.data_ptr(), dp.data_ptr(), n, d_out_affine=out.data_ptr())
    torch.cuda.synchronize()
    t = oracle.synth_point_scalars(78, 0, n)
    assert bytes(out.cpu().numpy()) == oracle.msm_expected_from_dlogs(ds.cpu().numpy(), t, n)


def test_headline_size_device_and_host_paths_checksum():
    """2^24 terms (BASELINE.json's headline size) with the library's own choices: the device-resident path picks the batched-affine
    kernel, the host entry point runs its term-chunk pipeline in which small chunks use XYZZ and large ones the affine kernel, all
    accumulating into the same buckets.  Both must equal the discrete-log checksum [sum s_i t_i] G."""
    import torch
    L = sv.CudaLoader(0)
    try:
        n = 1 << 24
        ds = torch.empty(n * 32, dtype=torch.uint8, device="cuda")
        dp = torch.empty(n * 64, dtype=torch.uint8, device="cuda")
        out = torch.zeros(64, dtype=torch.uint8, device="cuda")
        L.synth_scalars_device(91, 0, n, ds.data_ptr())
        L.synth_points_device(91, 0, n, dp.data_ptr())
        L.profile(True)
        L.msm_device(ds.data_ptr(), dp.data_ptr(), n, d_out_affine=out.data_ptr())
        torch.cuda.synchronize()
        assert any(name in ("msm_bucket_accumulate_affine", "msm_bucket_accumulate_chain") for name, _, _ in L.stage_times())
        L.profile(False)
        hs, hp = ds.cpu().numpy(), dp.cpu().numpy()
        exp = oracle.msm_expected_from_dlogs(hs, oracle.synth_point_scalars(91, 0, n), n)
        assert bytes(out.cpu().numpy()) == exp
        assert L.msm(hs, hp, n) == exp
    finally:
        L.close()


@pytest.mark.parametrize("big_mode", [2, 4], ids=["affine", "chain"])
def test_heavily_skewed_large(loader, big_mode):
    """2^20 terms with scalar 1: one bucket holds every term -> thousands of full-length tasks, all levels of the tree."""
    import torch
    loader.set_accumulate_mode(big_mode)
    try:
        n = 1 << 20
        dp = torch.empty(n * 64, dtype=torch.uint8, device="cuda")
        loader.synth_points_device(67, 0, n, dp.data_ptr())
        torch.cuda.synchronize()
        pts = dp.cpu().numpy()
        t = oracle.synth_point_scalars(67, 0, n)
        s = np.frombuffer(le(1) * n, dtype=np.uint8)
        assert loader.msm(s, pts, n) == oracle.msm_expected_from_dlogs(s, t, n)
    finally:
        loader.set_accumulate_mode(0)


@pytest.mark.parametrize("r", [8, 12, 16])
def test_chain_kernel_every_chain_count(r, monkeypatch):
    """SNARKV_BC_R selects the number of running sums per lane (the template instances of k_bucket_accumulate_chain)."""
    monkeypatch.setenv("SNARKV_BC_R", str(r))
    L = sv.CudaLoader(0)
    try:
        L.set_accumulate_mode(5)
        for n, c in ((3000, 4), (3000, 7), (1 << 14, 0), (50, 2)):
            s = oracle.synth_scalars(36, 0, n)
            p = oracle.synth_points(36, 0, n, 8)
            L.set_window_bits(c)
            assert L.msm(s, p, n) == oracle.msm_pippenger(s, p, n, 8), (r, n, c)
    finally:
        L.close()
