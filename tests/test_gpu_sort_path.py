"""GPU: the large-input sort path (csrc/sort.cuh: digits + coarse partition + per-partition bucket sort, no per-element global
atomics) must give the same MSM as the small-input path and the CPU oracle — uniform inputs, every window size it accepts, GLV on
and off, and the skewed digit distributions that overflow one partition / one bucket (all scalars equal, `Msm::base` scalars of 1,
two-valued scalars).  Bar: bit-exact."""
import numpy as np
import pytest
import torch

import oracle
import snark_verifier_b200 as sv
from oracle import bn254_model as m

pytestmark = pytest.mark.gpu
le = m.fe_to_le
N = (1 << 18) + 12345          # the two-level path starts at 2^18 (virtual) terms


@pytest.fixture(scope="module")
def loader():
    L = sv.CudaLoader(0)
    yield L
    L.close()


@pytest.fixture(scope="module")
def workload():
    s = oracle.synth_scalars(91, 0, N)
    p = oracle.synth_points(91, 0, N, 8)
    return s, p, oracle.msm_pippenger(s, p, N, 8)


@pytest.mark.parametrize("glv", [1, 2])
@pytest.mark.parametrize("c", [0, 9, 10, 13, 15, 16, 17, 19])
def test_sorted_path_matches_oracle_for_every_window(loader, workload, c, glv):
    s, p, exp = workload
    loader.set_window_bits(c)
    loader.set_glv_mode(glv)
    try:
        assert loader.msm(s, p, N) == exp, (c, glv)
    finally:
        loader.set_window_bits(0)
        loader.set_glv_mode(0)


@pytest.mark.parametrize("mode", [1, 2])
def test_sorted_path_both_accumulate_kernels(loader, workload, mode):
    s, p, exp = workload
    loader.set_accumulate_mode(mode)
    try:
        assert loader.msm(s, p, N) == exp
    finally:
        loader.set_accumulate_mode(0)


@pytest.mark.parametrize("val", [1, 2, 0xFFFF, (1 << 127) + 5, m.R - 1])
def test_sorted_path_all_scalars_equal(loader, val):
    """every term in ONE bucket of each window: one partition holds everything, one bucket exceeds the shared-memory buffer"""
    n = 1 << 18
    p = oracle.synth_points(92, 0, n, 8)
    s = le(val) * n
    # sum_i val * P_i = val * sum_i P_i : the oracle's Pippenger handles it too (single bucket per window)
    assert loader.msm(s, p, n) == oracle.msm_pippenger(s, p, n, 8), val


def test_sorted_path_two_valued_and_sparse_scalars(loader):
    n = 1 << 18
    rng = np.random.default_rng(7)
    p = oracle.synth_points(93, 0, n, 8)
    a, b = le(0x1234567 << 100), le(m.R - 77)
    pick = rng.integers(0, 2, size=n)
    s = b"".join(a if x else b for x in pick)
    assert loader.msm(s, p, n) == oracle.msm_pippenger(s, p, n, 8)
    # mostly-zero scalars: most digits are 0 and never enter the sort
    z = bytearray(32 * n)
    for i in rng.integers(0, n, size=5000):
        z[32 * i:32 * i + 32] = le(int(rng.integers(1, 1 << 62)) * 0x10001 % m.R)
    assert loader.msm(bytes(z), p, n) == oracle.msm_pippenger(bytes(z), p, n, 8)


def test_sorted_path_device_entry_and_host_chunk_pipeline_agree(loader):
    """2^22 terms: the host entry point cuts the input into term-chunks that are sorted separately into the same buckets"""
    n = 1 << 22
    ds = torch.empty(n * 32, dtype=torch.uint8, device="cuda")
    dp = torch.empty(n * 64, dtype=torch.uint8, device="cuda")
    out = torch.zeros(64, dtype=torch.uint8, device="cuda")
    loader.synth_scalars_device(94, 0, n, ds.data_ptr())
    loader.synth_points_device(94, 0, n, dp.data_ptr())
    loader.msm_device(ds.data_ptr(), dp.data_ptr(), n, d_out_affine=out.data_ptr())
    torch.cuda.synchronize()
    got_dev = bytes(out.cpu().numpy())
    s_host = ds.cpu().numpy()
    assert got_dev == oracle.msm_expected_from_dlogs(s_host, oracle.synth_point_scalars(94, 0, n), n)
    assert loader.msm(s_host, dp.cpu().numpy(), n) == got_dev
