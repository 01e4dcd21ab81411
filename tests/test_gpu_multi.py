"""GPU: the multi-device entry points (snarkv_multi_*, include/snarkv_cuda.h) and the resident-base MSM against the single-device
path and the CPU oracle.  The multi-device context is exercised on whatever the box has: with one GPU the same device is listed
twice (two contexts, two host threads, partials folded by the peer-memory kernel); with >= 2 GPUs real peer access is used as
well.  Bar: bit-exact."""
import numpy as np
import pytest
import torch

import oracle
import snark_verifier_b200 as sv
from oracle import bn254_model as m

pytestmark = pytest.mark.gpu
le = m.fe_to_le


def device_sets():
    n = torch.cuda.device_count()
    sets = [[0], [0, 0], [0, 0, 0]]
    if n >= 2:
        sets += [[0, 1], list(range(min(n, 8)))]
    return sets


@pytest.fixture(scope="module")
def loader():
    L = sv.CudaLoader(0)
    yield L
    L.close()


@pytest.mark.parametrize("devices", device_sets(), ids=lambda d: "dev" + "".join(map(str, d)))
def test_multi_msm_matches_single_device_and_oracle(loader, devices):
    M = sv.MultiCudaLoader(devices)
    try:
        assert M.n_devices == len(devices)
        for n in (1, 2, 3, 5, 1000, 1 << 14):
            s, p = oracle.synth_scalars(11, 0, n), oracle.synth_points(11, 0, n, 4)
            got = M.msm(s, p, n, flags=sv.CHECK_INPUTS)
            assert got == loader.msm(s, p, n), (devices, n)
            assert got == oracle.msm_pippenger(s, p, n, 4), (devices, n)
        with pytest.raises(sv.Error):
            M.msm(b"", b"", 0)                                   # empty slice: native.rs:69 panics
        bad = bytearray(oracle.synth_points(11, 0, 8, 1)); bad[64 * 6] ^= 1
        with pytest.raises(sv.Error):
            M.msm(oracle.synth_scalars(11, 0, 8), bytes(bad), 8, flags=sv.CHECK_INPUTS)   # invalid point on the LAST device's chunk
    finally:
        M.close()


@pytest.mark.parametrize("devices", device_sets()[1:], ids=lambda d: "dev" + "".join(map(str, d)))
def test_multi_decide_batch_shards_independent_checks(devices):
    g2 = oracle.g2_generator()
    sk = 0xABCDEF12345
    s_g2 = oracle.g2_mul(g2, le(sk))
    gen = m.g1_to_bytes(m.G1_GEN)
    n = 37
    a = [3 + 5 * i for i in range(n)]
    lhs = b"".join(oracle.g1_mul(gen, le(x * sk + (1 if i % 5 == 3 else 0))) for i, x in enumerate(a))   # every 5th is invalid
    rhs = b"".join(oracle.g1_mul(gen, le(x)) for x in a)
    M = sv.MultiCudaLoader(devices)
    try:
        M.set_deciding_key(sv.KzgDecidingKey(gen, g2, s_g2))
        acc, gt = M.decide_batch(lhs, rhs, n, want_gt=True)
        oacc, ogt = oracle.kzg_decide_batch(lhs, rhs, n, g2, s_g2, 4, want_gt=True)
        assert acc == oacc and gt == ogt
        assert acc == bytes(0 if i % 5 == 3 else 1 for i in range(n))
    finally:
        M.close()


@pytest.mark.parametrize("devices", device_sets()[1:], ids=lambda d: "dev" + "".join(map(str, d)))
def test_multi_msm_batch_rlc_matches_single_device(loader, devices):
    rng = np.random.default_rng(3)
    sizes = [int(x) for x in rng.integers(1, 25, size=41)]
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
    total = int(offs[-1])
    s, p = oracle.synth_scalars(21, 0, total), oracle.synth_points(21, 0, total, 4)
    rho = le(0x1234567890ABCDEF1122334455667788)
    expect = loader.msm_batch_rlc(s, p, offs, rho, flags=sv.CHECK_INPUTS)
    # independent model: sum_j rho^j MSM_j with the oracle
    acc, rj = bytes(64), 1
    for j, sz in enumerate(sizes):
        lo = int(offs[j])
        part = oracle.msm_pippenger(s[32 * lo:32 * (lo + sz)], p[64 * lo:64 * (lo + sz)], sz, 1)
        acc = oracle.g1_add(acc, oracle.g1_mul(part, le(rj)))
        rj = rj * int.from_bytes(rho, "little") % m.R
    assert expect == acc
    M = sv.MultiCudaLoader(devices)
    try:
        assert M.msm_batch_rlc(s, p, offs, rho, flags=sv.CHECK_INPUTS) == expect
    finally:
        M.close()


def test_msm_with_resident_bases_matches_plain_msm(loader):
    for n in (1, 5, 300, 1 << 14, (1 << 16) + 7):
        s, p = oracle.synth_scalars(31, 0, n), oracle.synth_points(31, 0, n, 4)
        h = loader.bases_upload(p, n)
        try:
            for seed in (31, 32):                        # the base set outlives many scalar vectors
                s = oracle.synth_scalars(seed, 0, n)
                assert loader.msm_bases_resident(h, s, n) == loader.msm(s, p, n), n
            with pytest.raises(sv.Error):
                loader.msm_bases_resident(h, s, n - 1 if n > 1 else 2)   # one scalar per resident base
        finally:
            loader.bases_free(h)
    bad = bytearray(oracle.synth_points(31, 0, 4, 1)); bad[3] ^= 0x40
    with pytest.raises(sv.Error):
        loader.bases_upload(bytes(bad), 4)               # validated at upload (from_xy semantics)


def test_msm_with_resident_bases_montgomery_format():
    L = sv.CudaLoader(0, fmt=sv.MONTGOMERY)
    try:
        n = 3000
        s, p = oracle.synth_scalars(41, 0, n), oracle.synth_points(41, 0, n, 4)
        sm, pm = oracle.to_mont_batch(1, s, n, 2), oracle.to_mont_batch(0, p, 2 * n, 2)
        h = L.bases_upload(pm, n)
        got = L.msm_bases_resident(h, sm, n)
        L.bases_free(h)
        assert got == L.msm(sm, pm, n)
        assert got == oracle.to_mont_batch(0, oracle.msm_pippenger(s, p, n, 4), 2, 1)
    finally:
        L.close()
