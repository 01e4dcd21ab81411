"""GPU parity suite: the CUDA path (through the C-ABI of include/snarkv_cuda.h) against
  (1) the committed golden vectors (tests/golden, from the independent Python big-int model),
  (2) the C++ oracle (oracle/) on seeded inputs at sizes it finishes in seconds,
  (3) size-independent properties at full size (discrete-log checksum, linearity).
Bar: bit-exact (integer arithmetic only).  Shapes follow the reference's own tests: accept valid, reject tampered
(system/halo2/test/kzg/native.rs:57-68, test/kzg/evm.rs:58-62), mock accumulator (s*G, G) (test/kzg.rs:37-45)."""
import ctypes

import numpy as np
import pytest

import oracle
import snark_verifier_b200 as sv
from oracle import bn254_model as m

pytestmark = pytest.mark.gpu
H = bytes.fromhex
le = m.fe_to_le


@pytest.fixture(scope="module")
def loader():
    L = sv.CudaLoader(0)
    yield L
    L.close()


@pytest.fixture(scope="module")
def mont_loader():
    L = sv.CudaLoader(0, fmt=sv.MONTGOMERY)
    yield L
    L.close()


def to_mont_pts(pb):
    return b"".join(oracle.fp_to_mont(0, pb[i:i + 32]) for i in range(0, len(pb), 32))


def to_mont_scalars(sb):
    return b"".join(oracle.fp_to_mont(1, sb[i:i + 32]) for i in range(0, len(sb), 32))


# ---------------------------------------------------------------------------------------------------------------------
# field layer
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("field,name", [(0, "fq"), (1, "fr")])
def test_device_field_ops_match_golden(loader, golden, field, name):
    g = golden("field")[name]
    a = b"".join(H(x[0]) for x in g["mul"]); b = b"".join(H(x[1]) for x in g["mul"])
    n = len(g["mul"])
    assert loader.field_op(field, 0, a, b, n) == b"".join(H(x[2]) for x in g["mul"])
    mod = m.P if field == 0 else m.R
    ai = [int.from_bytes(H(x[0]), "little") for x in g["mul"]]
    bi = [int.from_bytes(H(x[1]), "little") for x in g["mul"]]
    assert loader.field_op(field, 1, a, b, n) == b"".join(le((x + y) % mod) for x, y in zip(ai, bi))
    assert loader.field_op(field, 2, a, b, n) == b"".join(le((x - y) % mod) for x, y in zip(ai, bi))
    inv_in = b"".join(H(x[0]) for x in g["inv"])
    assert loader.field_op(field, 3, inv_in, inv_in, len(g["inv"])) == b"".join(H(x[1]) for x in g["inv"])


def test_device_field_mul_random_vs_python(loader):
    rng = np.random.default_rng(5)
    n = 4096
    for field, mod in ((0, m.P), (1, m.R)):
        xs = [int.from_bytes(rng.bytes(32), "little") % mod for _ in range(n)]
        ys = [int.from_bytes(rng.bytes(32), "little") % mod for _ in range(n)]
        got = loader.field_op(field, 0, b"".join(map(le, xs)), b"".join(map(le, ys)), n)
        assert got == b"".join(le(x * y % mod) for x, y in zip(xs, ys))


# ---------------------------------------------------------------------------------------------------------------------
# MSM
# ---------------------------------------------------------------------------------------------------------------------
def test_msm_golden_cases(loader, golden):
    for case in golden("msm"):
        got = loader.msm(H(case["scalars"]), H(case["points"]), case["n"], flags=sv.CHECK_INPUTS)
        assert got == H(case["expected"]), case["name"]


def test_msm_golden_cases_montgomery_layout(mont_loader, golden):
    """halo2curves in-memory layout in, same layout out (the zero-copy path of the Rust glue)."""
    for case in golden("msm"):
        got = mont_loader.msm(to_mont_scalars(H(case["scalars"])), to_mont_pts(H(case["points"])), case["n"])
        assert got == to_mont_pts(H(case["expected"])), case["name"]


def test_msm_pairs_api_is_native_loader_shape(loader, golden):
    case = golden("msm")[3]
    s, p, n = H(case["scalars"]), H(case["points"]), case["n"]
    pairs = [(s[32 * i:32 * i + 32], p[64 * i:64 * i + 64]) for i in range(n)]
    assert loader.multi_scalar_multiplication(pairs) == H(case["expected"])


@pytest.mark.parametrize("n", [1, 2, 5, 33, 100, 1000, 4097, 1 << 14])
def test_msm_vs_oracle_seeded(loader, n):
    s = oracle.synth_scalars(21, 0, n)
    p = oracle.synth_points(21, 0, n, 8)
    exp = oracle.msm_pippenger(s, p, n, 8)
    assert loader.msm(s, p, n) == exp
    if n <= 100:
        assert exp == oracle.msm_native(s, p, n)  # the literal NativeLoader fold (native.rs:61-71)


@pytest.mark.parametrize("c", [2, 3, 4, 8, 11, 13, 16])
def test_msm_every_window_size_agrees(loader, c):
    n = 3000
    s = oracle.synth_scalars(22, 0, n)
    p = oracle.synth_points(22, 0, n, 8)
    exp = oracle.msm_pippenger(s, p, n, 8)
    loader.set_window_bits(c)
    try:
        assert loader.msm(s, p, n) == exp
    finally:
        loader.set_window_bits(0)


def test_msm_skewed_scalars_single_bucket(loader):
    """All scalars equal / tiny scalars: every term lands in the same bucket of each window (worst-case balance)."""
    n = 5000
    p = oracle.synth_points(23, 0, n, 8)
    for val in (1, 2, 0xFFFF, m.R - 1):
        s = le(val) * n
        assert loader.msm(s, p, n) == oracle.msm_pippenger(s, p, n, 8), val


def test_msm_empty_and_invalid_inputs(loader):
    with pytest.raises(sv.Error):
        loader.msm(b"", b"", 0)                      # native.rs:69 .unwrap() on empty
    g = m.g1_to_bytes(m.G1_GEN)
    with pytest.raises(sv.Error):
        loader.msm(le(m.R), g, 1, flags=sv.CHECK_INPUTS)      # non-canonical scalar: from_repr fails
    off = le(1) + le(3)
    with pytest.raises(sv.Error):
        loader.msm(le(5), off, 1, flags=sv.CHECK_INPUTS)      # not on the curve: from_xy fails
    with pytest.raises(sv.Error):
        loader.msm(le(5), le(m.P) + le(2), 1, flags=sv.CHECK_INPUTS)


def test_msm_batch_matches_native_fold(loader):
    sizes = [1, 21, 3, 20, 1, 24, 7]              # StandardPlonk shapes: 21+3 (GWC), 20+1 (SHPLONK), SURVEY §3.1
    offs = [0]
    for k in sizes:
        offs.append(offs[-1] + k)
    total = offs[-1]
    s = oracle.synth_scalars(31, 0, total)
    p = oracle.synth_points(31, 0, total, 8)
    got = loader.msm_batch(s, p, offs)
    for j, k in enumerate(sizes):
        lo = offs[j]
        assert got[j] == oracle.msm_native(s[32 * lo:32 * (lo + k)], p[64 * lo:64 * (lo + k)], k), j


def test_external_known_answer_on_device(loader):
    """EIP-196 `bn256Add` vector (public Ethereum precompile tests; BN254 = alt_bn128): (1, 2) + (1, 2), straight from the device."""
    two_g1 = (0x030644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD3,
              0x15ED738C0E0A7C92E7845F96B2AE9C0A68A6A449E3538FC7FF3EBF7A5A18A2C4)
    g = m.g1_to_bytes(m.G1_GEN)
    exp = le(two_g1[0]) + le(two_g1[1])
    assert loader.msm(le(1) * 2, g * 2, 2) == exp
    assert loader.msm(le(2), g, 1) == exp


def test_msm_host_mirror_evaluate(loader):
    """util::msm::Msm algebra on the host + evaluate(Some(gen)) through the loader."""
    pts = [oracle.synth_points(41, i, 1) for i in range(3)]
    a = sv.Msm.base(loader, pts[0]) * 5 + sv.Msm.base(loader, pts[1]) * 7 + sv.Msm.base(loader, pts[0]) * 11
    a = a + sv.Msm.constant_(loader, 13)
    assert len(a.bases) == 2                                   # push dedupes equal bases (util/msm.rs:109-116)
    g = m.g1_to_bytes(m.G1_GEN)
    exp = oracle.msm_native(le(13) + le(16) + le(7), g + pts[0] + pts[1], 3)
    assert a.evaluate(g) == exp


# ---------------------------------------------------------------------------------------------------------------------
# device-resident operands, synthetic generators, Jacobian partials
# ---------------------------------------------------------------------------------------------------------------------
def test_synth_generators_match_oracle(loader):
    import torch
    n = 3000
    ds = torch.empty(n * 32, dtype=torch.uint8, device="cuda")
    dp = torch.empty(n * 64, dtype=torch.uint8, device="cuda")
    loader.synth_scalars_device(9, 100, n, ds.data_ptr())
    loader.synth_points_device(9, 100, n, dp.data_ptr())
    torch.cuda.synchronize()
    assert bytes(ds.cpu().numpy()) == oracle.synth_scalars(9, 100, n)
    assert bytes(dp.cpu().numpy()) == oracle.synth_points(9, 100, n, 8)


def test_device_msm_partials_fold_like_rayon_chunks(loader):
    """util/msm.rs:322-336: chunk the terms, one partial per chunk, fold the Jacobian partials, to_affine."""
    import torch
    n, k = 6000, 3
    s = oracle.synth_scalars(12, 0, n); p = oracle.synth_points(12, 0, n, 8)
    ds = torch.frombuffer(bytearray(s), dtype=torch.uint8).cuda()
    dp = torch.frombuffer(bytearray(p), dtype=torch.uint8).cuda()
    parts = torch.zeros(k * 96, dtype=torch.uint8, device="cuda")
    out = torch.zeros(64, dtype=torch.uint8, device="cuda")
    chunk = n // k
    for j in range(k):
        loader.msm_device(ds.data_ptr() + 32 * j * chunk, dp.data_ptr() + 64 * j * chunk, chunk,
                          d_out_jacobian=parts.data_ptr() + 96 * j)
        torch.cuda.synchronize()
    loader.fold_partials_device(parts.data_ptr(), k, out.data_ptr())
    torch.cuda.synchronize()
    assert bytes(out.cpu().numpy()) == oracle.msm_pippenger(s, p, n, 8)


@pytest.mark.parametrize("logn", [20, 22])
def test_msm_full_size_dlog_checksum(loader, logn):
    """Size-independent property at BASELINE sizes: P_i = [t_i]G  =>  MSM = [sum s_i t_i mod r] G."""
    import torch
    n = 1 << logn
    ds = torch.empty(n * 32, dtype=torch.uint8, device="cuda")
    dp = torch.empty(n * 64, dtype=torch.uint8, device="cuda")
    out = torch.zeros(64, dtype=torch.uint8, device="cuda")
    loader.synth_scalars_device(77, 0, n, ds.data_ptr())
    loader.synth_points_device(77, 0, n, dp.data_ptr())
    loader.msm_device(ds.data_ptr(), dp.data_ptr(), n, d_out_affine=out.data_ptr())
    torch.cuda.synchronize()
    s_host = ds.cpu().numpy()
    t = oracle.synth_point_scalars(77, 0, n)
    assert bytes(s_host[:32 * 64]) == oracle.synth_scalars(77, 0, 64)
    exp = oracle.msm_expected_from_dlogs(s_host, t, n)
    assert bytes(out.cpu().numpy()) == exp
    # spot-check generated points against the oracle's own scalar multiplication
    pts = dp[: 64 * 16].cpu().numpy().tobytes()
    assert pts == oracle.synth_points(77, 0, 16, 1)


def test_msm_partial_host_entry(loader):
    import torch
    n = 3000
    s = oracle.synth_scalars(13, 0, n); p = oracle.synth_points(13, 0, n, 8)
    part = torch.zeros(96, dtype=torch.uint8, device="cuda")
    out = torch.zeros(64, dtype=torch.uint8, device="cuda")
    loader.msm_partial(s, p, n, part.data_ptr())
    loader.fold_partials_device(part.data_ptr(), 1, out.data_ptr())
    torch.cuda.synchronize()
    assert bytes(out.cpu().numpy()) == oracle.msm_pippenger(s, p, n, 8)


def test_msm_linearity_property(loader):
    """MSM(s, P) + MSM(s', P) == MSM(s + s', P) at a size the oracle never sees."""
    import torch
    n = 1 << 18
    dp = torch.empty(n * 64, dtype=torch.uint8, device="cuda")
    loader.synth_points_device(55, 0, n, dp.data_ptr())
    torch.cuda.synchronize()
    s1 = np.frombuffer(oracle.synth_scalars(55, 0, n), dtype=np.uint8)
    s2 = np.frombuffer(oracle.synth_scalars(56, 0, n), dtype=np.uint8)
    a = [int.from_bytes(s1[32 * i:32 * i + 32].tobytes(), "little") for i in range(n)]
    b = [int.from_bytes(s2[32 * i:32 * i + 32].tobytes(), "little") for i in range(n)]
    s3 = b"".join(le((x + y) % m.R) for x, y in zip(a, b))
    pts = dp.cpu().numpy().tobytes()
    r1, r2, r3 = loader.msm(s1.tobytes(), pts, n), loader.msm(s2.tobytes(), pts, n), loader.msm(s3, pts, n)
    assert oracle.g1_add(r1, r2) == r3


# ---------------------------------------------------------------------------------------------------------------------
# KZG decide / accumulate
# ---------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def kzg(loader, golden):
    g = golden("pairing")
    dk = sv.KzgDecidingKey(m.g1_to_bytes(m.G1_GEN), H(g["g2_generator"]), H(g["s_g2"]))
    return sv.KzgAs(loader, dk)


@pytest.fixture(params=[1, 3, 4, 5], ids=["thread_per_check", "block_per_check", "warp_per_check", "latency_kernel"])
def pairing_mode(request, loader):
    loader.set_pairing_mode(request.param)
    yield request.param
    loader.set_pairing_mode(0)


def test_decide_golden_accept_reject_and_gt_bytes(kzg, golden, pairing_mode):
    g = golden("pairing")
    checks = g["checks"]
    lhs = b"".join(H(c["lhs"]) for c in checks); rhs = b"".join(H(c["rhs"]) for c in checks)
    acc, gt = kzg.decide_batch(lhs, rhs, len(checks), want_gt=True)
    assert list(acc) == [int(c["accept"]) for c in checks]
    for i, c in enumerate(checks):
        assert gt[384 * i:384 * (i + 1)] == H(c["gt"]), c["name"]
    # e(G1, G2) itself: lhs = G, rhs = identity
    acc, gt = kzg.decide_batch(m.g1_to_bytes(m.G1_GEN), bytes(64), 1, want_gt=True)
    assert acc == b"\x00" and gt == H(g["e_G1_G2"])


def test_decide_and_decide_all_error_behaviour(kzg, golden):
    checks = {c["name"]: c for c in golden("pairing")["checks"]}
    ok = sv.KzgAccumulator(H(checks["mock_sG_G"]["lhs"]), H(checks["mock_sG_G"]["rhs"]))
    bad = sv.KzgAccumulator(H(checks["tampered_rhs"]["lhs"]), H(checks["tampered_rhs"]["rhs"]))
    kzg.decide(ok)
    kzg.decide_all([ok, ok])
    kzg.decide_all([])
    with pytest.raises(sv.AssertionFailure, match="e\\(lhs, g2\\)"):
        kzg.decide(bad)
    with pytest.raises(sv.AssertionFailure):
        kzg.decide_all([ok, bad, ok])


def test_decide_batch_vs_oracle_random_mix(kzg, golden, pairing_mode):
    g = golden("pairing")
    s = int.from_bytes(H(g["s"]), "little")
    rng = np.random.default_rng(8)
    n = 96
    gen = m.g1_to_bytes(m.G1_GEN)
    lhs, rhs = [], []
    for i in range(n):
        a = int.from_bytes(rng.bytes(31), "little") + 1
        bad = (i % 3 == 1)
        lhs.append(oracle.g1_mul(gen, le((a * s + (1 if bad else 0)) % m.R)))
        rhs.append(oracle.g1_mul(gen, le(a)))
    lhs, rhs = b"".join(lhs), b"".join(rhs)
    acc, gt = kzg.decide_batch(lhs, rhs, n, want_gt=True)
    oacc, ogt = oracle.kzg_decide_batch(lhs, rhs, n, H(g["g2_generator"]), H(g["s_g2"]), threads=8, hoist=0, want_gt=True)
    assert acc == oacc and gt == ogt
    assert list(acc) == [0 if i % 3 == 1 else 1 for i in range(n)]


def test_decide_rejects_off_curve_accumulator(kzg, pairing_mode):
    acc, _ = kzg.decide_batch(le(1) + le(3), m.g1_to_bytes(m.G1_GEN), 1)
    assert acc == b"\x00"


def test_latency_kernel_dead_pairs_and_grid_stride(kzg, golden, loader):
    """pairing_fast.cu against the thread-per-check kernels: identity on either side (the merged-line tables have one row set per
    live-pair case), both sides identity, and more checks than resident blocks (grid-stride reuse of the shared-memory state)."""
    g = golden("pairing")
    s = int.from_bytes(H(g["s"]), "little")
    gen = m.g1_to_bytes(m.G1_GEN)
    n = 700
    lhs, rhs = [], []
    for i in range(n):
        a = 1000003 * i + 17
        l = oracle.g1_mul(gen, le((a * s + (1 if i % 5 == 2 else 0)) % m.R))
        r = oracle.g1_mul(gen, le(a))
        if i % 7 == 3: l = bytes(64)
        if i % 11 == 4: r = bytes(64)
        lhs.append(l); rhs.append(r)
    lhs, rhs = b"".join(lhs), b"".join(rhs)
    loader.set_pairing_mode(1)
    acc1, gt1 = kzg.decide_batch(lhs, rhs, n, want_gt=True)
    loader.set_pairing_mode(5)
    acc5, gt5 = kzg.decide_batch(lhs, rhs, n, want_gt=True)
    loader.set_pairing_mode(0)
    assert acc5 == acc1 and gt5 == gt1
    both = [i for i in range(n) if i % 7 == 3 and i % 11 == 4]
    assert both and all(acc5[i] == 1 for i in both)            # e(O, .) e(O, .) = 1
    oacc, ogt = oracle.kzg_decide_batch(lhs[:64 * 48], rhs[:64 * 48], 48, H(g["g2_generator"]), H(g["s_g2"]), threads=8, hoist=0, want_gt=True)
    assert acc5[:48] == oacc and gt5[:384 * 48] == ogt


def test_decide_needs_key():
    L = sv.CudaLoader(0)
    try:
        acc = ctypes.create_string_buffer(1)
        rc = L.lib.snarkv_kzg_decide_batch(L.h, bytes(64), bytes(64), 1, sv.CANONICAL, acc, None)
        assert rc == sv.ERR_NO_KEY
    finally:
        L.close()


def test_bad_deciding_key_is_rejected():
    L = sv.CudaLoader(0)     # own context: a failed set leaves that context without a key
    try:
        with pytest.raises(sv.Error):
            sv.KzgAs(L, sv.KzgDecidingKey(bytes(64), le(1) + le(2) + le(3) + le(4), bytes(128)))
    finally:
        L.close()


def test_accumulate_golden_then_decide(kzg, golden):
    for c in golden("accumulate"):
        n = c["n"]
        lhs, rhs = H(c["lhs"]), H(c["rhs"])
        accs = [sv.KzgAccumulator(lhs[64 * i:64 * i + 64], rhs[64 * i:64 * i + 64]) for i in range(n)]
        out = kzg.verify(accs, H(c["r"]))
        assert out.lhs == H(c["out_lhs"]) and out.rhs == H(c["out_rhs"])
        kzg.decide(out)          # an accumulation of valid accumulators is valid


def test_accumulate_256_vs_oracle(kzg, golden):
    """BASELINE config 4 shape: 256 accumulators (a_i s G, a_i G), powers of r, then one decide."""
    g = golden("pairing")
    s = int.from_bytes(H(g["s"]), "little")
    n = 256
    gen = m.g1_to_bytes(m.G1_GEN)
    a = [int.from_bytes(oracle.synth_scalars(91, i, 1), "little") for i in range(n)]
    lhs = b"".join(oracle.g1_mul(gen, le(x * s % m.R)) for x in a)
    rhs = b"".join(oracle.g1_mul(gen, le(x)) for x in a)
    r = oracle.synth_scalars(92, 0, 1)
    accs = [sv.KzgAccumulator(lhs[64 * i:64 * i + 64], rhs[64 * i:64 * i + 64]) for i in range(n)]
    out = kzg.verify(accs, r)
    ol, orr = oracle.kzg_accumulate(lhs, rhs, n, r)
    assert (out.lhs, out.rhs) == (ol, orr)
    kzg.decide(out)


def test_cpp_host_mirror(tmp_path, golden):
    """Compile and run the C++ mirror of the reference's Loader/Decider surface (host/cuda_loader.hpp) against golden data."""
    import os, struct, subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "host_mirror_test"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-o", str(exe), os.path.join(root, "tests", "host_mirror_test.cpp"),
                           "-L" + os.path.join(root, "snark_verifier_b200"), "-lsnarkv_cuda",
                           "-Wl,-rpath," + os.path.join(root, "snark_verifier_b200")])
    case = golden("msm")[3]
    g = golden("pairing")
    checks = {c["name"]: c for c in g["checks"]}
    blob = struct.pack("<I", case["n"]) + H(case["scalars"]) + H(case["points"]) + H(case["expected"])
    blob += H(g["g2_generator"]) + H(g["s_g2"])
    blob += H(checks["valid_0"]["lhs"]) + H(checks["valid_0"]["rhs"]) + H(checks["tampered_rhs"]["rhs"])
    # GWC19 section: the toy-prover fixture of test_pcs_boundary.py; expected accumulator = the Python mirror over the NativeLoader fold
    from snark_verifier_b200 import pcs
    from test_pcs_boundary import GEN, OracleNativeLoader, make_fixture
    from test_pcs_mirror import gwc19_queries
    fx = make_fixture(seed=31, k=5)
    qs = gwc19_queries(fx)
    N = OracleNativeLoader()
    exp = pcs.Gwc19.verify(N, GEN, [sv.Msm.base(N, c) for c in fx["C"]], fx["points"][0], qs, pcs.Gwc19Proof(fx["v"], fx["W"], fx["u"]))
    blob += struct.pack("<II", len(fx["C"]), len(qs)) + le(fx["points"][0]) + le(fx["v"]) + le(fx["u"])
    for q in qs:
        blob += struct.pack("<I", q.poly) + le(q.shift) + le(q.eval)
    blob += b"".join(fx["C"]) + struct.pack("<I", len(fx["W"])) + b"".join(fx["W"])
    blob += fx["g2"] + fx["s_g2"] + GEN + exp.lhs + exp.rhs
    case8 = [c for c in golden("plonk_eval")["cases"] if c["k"] == 8][0]
    ins = [int(v, 16) for r in case8["rows"] for v in r["inputs"]]
    outs = [int(v, 16) for r in case8["rows"] for v in r["outputs"]]
    nrows = len(case8["rows"])
    blob += struct.pack("<III", nrows, len(ins) // nrows, len(outs) // nrows) + b"".join(le(v) for v in ins) + b"".join(le(v) for v in outs)
    inp = tmp_path / "in.bin"
    inp.write_bytes(blob)
    out = subprocess.run([str(exe), str(inp)], capture_output=True, text=True)
    assert out.returncode == 0 and "host mirror ok" in out.stdout, out.stderr


def test_fused_rlc_decide_all_accepts_valid_rejects_one_bad(kzg, golden):
    """decider.rs:146-185 shape: accumulate with powers of a challenge, one pairing.  4096 accumulators (BASELINE config 3)."""
    g = golden("pairing")
    s = int.from_bytes(H(g["s"]), "little")
    n = 4096
    gen = m.g1_to_bytes(m.G1_GEN)
    base_l = oracle.g1_mul(gen, le(s)); base_r = gen
    a = [int.from_bytes(oracle.synth_scalars(71, i, 1), "little") for i in range(64)]
    lhs64 = [oracle.g1_mul(base_l, le(x)) for x in a]
    rhs64 = [oracle.g1_mul(base_r, le(x)) for x in a]
    lhs = b"".join(lhs64[i % 64] for i in range(n)); rhs = b"".join(rhs64[i % 64] for i in range(n))
    rho = oracle.synth_scalars(72, 0, 1)
    ok, acc = kzg.decide_all_fused(lhs, rhs, n, rho)
    assert ok
    # the combined accumulator equals the oracle's accumulate on a prefix-sized instance
    k = 200
    ok2, acc2 = kzg.decide_all_fused(lhs[:64 * k], rhs[:64 * k], k, rho)
    assert ok2 and (acc2.lhs, acc2.rhs) == oracle.kzg_accumulate(lhs[:64 * k], rhs[:64 * k], k, rho)
    bad = bytearray(rhs); bad[64 * 1234:64 * 1235] = oracle.g1_mul(gen, le(a[1234 % 64] + 1))
    ok3, _ = kzg.decide_all_fused(lhs, bytes(bad), n, rho)
    assert not ok3
    off = bytearray(rhs); off[64 * 7:64 * 8] = le(1) + le(3)      # not on the curve
    ok4, _ = kzg.decide_all_fused(lhs, bytes(off), n, rho)
    assert not ok4


def test_host_entry_chunk_pipeline_matches_device_path(loader):
    """n >= 2^22 goes through the 4-chunk copy/compute pipeline of snarkv_g1_msm; result must equal the resident path and the
    discrete-log checksum.  n is not a multiple of 4 on purpose (ragged last chunk)."""
    import torch
    n = (1 << 22) + 3
    ds = torch.empty(n * 32, dtype=torch.uint8, device="cuda")
    dp = torch.empty(n * 64, dtype=torch.uint8, device="cuda")
    out = torch.zeros(64, dtype=torch.uint8, device="cuda")
    loader.synth_scalars_device(88, 0, n, ds.data_ptr())
    loader.synth_points_device(88, 0, n, dp.data_ptr())
    loader.msm_device(ds.data_ptr(), dp.data_ptr(), n, d_out_affine=out.data_ptr())
    torch.cuda.synchronize()
    hs, hp = ds.cpu().numpy(), dp.cpu().numpy()
    got = loader.msm(hs, hp, n)
    assert got == bytes(out.cpu().numpy())
    assert got == oracle.msm_expected_from_dlogs(hs, oracle.synth_point_scalars(88, 0, n), n)
    part = torch.zeros(96, dtype=torch.uint8, device="cuda")
    loader.msm_partial(hs, hp, n, part.data_ptr())
    loader.fold_partials_device(part.data_ptr(), 1, out.data_ptr())
    torch.cuda.synchronize()
    assert bytes(out.cpu().numpy()) == got
    # same pipeline on halo2curves' in-memory layout (no conversion kernel), with input validation on
    Lm = sv.CudaLoader(0, fmt=sv.MONTGOMERY)
    try:
        ms = oracle.to_mont_batch(1, hs, n, 8); mp = oracle.to_mont_batch(0, hp, 2 * n, 8)
        assert Lm.msm(ms, mp, n, flags=sv.CHECK_INPUTS) == to_mont_pts(got)
    finally:
        Lm.close()


def test_msm_batch_rlc_equals_combination_of_native_folds(loader):
    """sum_j rho^j * MSM_j computed as one fused MSM == the same combination of the per-proof NativeLoader folds."""
    sizes = [21, 3, 20, 1, 24, 7, 21, 3]
    offs = [0]
    for k in sizes:
        offs.append(offs[-1] + k)
    total = offs[-1]
    s = oracle.synth_scalars(61, 0, total); p = oracle.synth_points(61, 0, total, 8)
    rho = oracle.synth_scalars(62, 0, 1)
    got = loader.msm_batch_rlc(s, p, offs, rho, flags=sv.CHECK_INPUTS)
    r = int.from_bytes(rho, "little")
    acc = bytes(64)
    for j, k in enumerate(sizes):
        lo = offs[j]
        part = oracle.msm_native(s[32 * lo:32 * (lo + k)], p[64 * lo:64 * (lo + k)], k)
        acc = oracle.g1_add(acc, oracle.g1_mul(part, le(pow(r, j, m.R))))
    assert got == acc


def test_msm_heavily_skewed_large(loader):
    """2^20 terms, every scalar = 1 (Msm::base) or = r-1: one bucket per window holds every term -> 2048+ tasks per bucket."""
    import torch
    n = 1 << 20
    dp = torch.empty(n * 64, dtype=torch.uint8, device="cuda")
    loader.synth_points_device(66, 0, n, dp.data_ptr())
    torch.cuda.synchronize()          # the *_device entry points are asynchronous on the loader's own stream
    pts = dp.cpu().numpy()
    t = oracle.synth_point_scalars(66, 0, n)
    for val in (1, m.R - 1):
        s = np.frombuffer(le(val) * n, dtype=np.uint8)
        assert loader.msm(s, pts, n) == oracle.msm_expected_from_dlogs(s, t, n), val


# ---------------------------------------------------------------------------------------------------------------------
# halo2curves in-memory layout (SNARKV_MONTGOMERY) through the KZG entry points, device-variant validation, odd inputs
# ---------------------------------------------------------------------------------------------------------------------
def test_decide_and_accumulate_in_montgomery_layout(mont_loader, golden):
    g = golden("pairing")
    kz = sv.KzgAs(mont_loader, sv.KzgDecidingKey(m.g1_to_bytes(m.G1_GEN), H(g["g2_generator"]), H(g["s_g2"])))   # key is always canonical
    checks = g["checks"]
    lhs = to_mont_pts(b"".join(H(c["lhs"]) for c in checks)); rhs = to_mont_pts(b"".join(H(c["rhs"]) for c in checks))
    for mode in (1, 3, 4, 5):
        mont_loader.set_pairing_mode(mode)
        acc, gt = kz.decide_batch(lhs, rhs, len(checks), want_gt=True)
        assert list(acc) == [int(c["accept"]) for c in checks]
        assert gt == b"".join(H(c["gt"]) for c in checks)          # GT is always canonical bytes
    mont_loader.set_pairing_mode(0)
    c = golden("accumulate")[2]
    n = c["n"]
    ml, mr = to_mont_pts(H(c["lhs"])), to_mont_pts(H(c["rhs"]))
    accs = [sv.KzgAccumulator(ml[64 * i:64 * i + 64], mr[64 * i:64 * i + 64]) for i in range(n)]
    out = kz.verify(accs, oracle.fp_to_mont(1, H(c["r"])))
    assert (out.lhs, out.rhs) == (to_mont_pts(H(c["out_lhs"])), to_mont_pts(H(c["out_rhs"])))


def test_device_entry_validation_and_alignment(loader):
    import torch
    n = 64
    s = bytearray(oracle.synth_scalars(5, 0, n)); p = bytearray(oracle.synth_points(5, 0, n, 2))
    s[32 * 7:32 * 8] = le(m.R)                      # non-canonical scalar
    ds = torch.frombuffer(s, dtype=torch.uint8).cuda(); dp = torch.frombuffer(p, dtype=torch.uint8).cuda()
    out = torch.zeros(64, dtype=torch.uint8, device="cuda"); st = torch.zeros(4, dtype=torch.uint8, device="cuda")
    loader.msm_device(ds.data_ptr(), dp.data_ptr(), n, d_out_affine=out.data_ptr(), d_status=st.data_ptr(), flags=sv.CHECK_INPUTS)
    torch.cuda.synchronize()
    assert int.from_bytes(bytes(st.cpu().numpy()), "little", signed=True) == sv.ERR_BAD_SCALAR
    p[64 * 9 + 32] ^= 1                              # y of point 9 off the curve
    s[32 * 7:32 * 8] = le(5)
    ds = torch.frombuffer(s, dtype=torch.uint8).cuda(); dp = torch.frombuffer(p, dtype=torch.uint8).cuda()
    loader.msm_device(ds.data_ptr(), dp.data_ptr(), n, d_out_affine=out.data_ptr(), d_status=st.data_ptr(), flags=sv.CHECK_INPUTS)
    torch.cuda.synchronize()
    assert int.from_bytes(bytes(st.cpu().numpy()), "little", signed=True) == sv.ERR_BAD_POINT
    with pytest.raises(sv.Error):                    # the TMA-staged scalar stream needs 16-byte alignment
        loader.msm_device(ds.data_ptr() + 4, dp.data_ptr(), n - 1, d_out_affine=out.data_ptr())


def test_fold_partials_with_identity_and_cancelling_partials(loader):
    import torch
    gen = m.g1_to_bytes(m.G1_GEN)
    parts = torch.zeros(4 * 96, dtype=torch.uint8, device="cuda")
    out = torch.zeros(64, dtype=torch.uint8, device="cuda")
    ds = torch.frombuffer(bytearray(le(7) + le(m.R - 7) + le(0) + le(3)), dtype=torch.uint8).cuda()
    dp = torch.frombuffer(bytearray(gen * 4), dtype=torch.uint8).cuda()
    for j in range(4):                               # partials: 7G, -7G, identity, 3G
        loader.msm_device(ds.data_ptr() + 32 * j, dp.data_ptr() + 64 * j, 1, d_out_jacobian=parts.data_ptr() + 96 * j)
        torch.cuda.synchronize()
    loader.fold_partials_device(parts.data_ptr(), 4, out.data_ptr()); torch.cuda.synchronize()
    assert bytes(out.cpu().numpy()) == oracle.g1_mul(gen, le(3))
    loader.fold_partials_device(parts.data_ptr(), 3, out.data_ptr()); torch.cuda.synchronize()
    assert bytes(out.cpu().numpy()) == bytes(64)     # 7G - 7G + O = identity = (0, 0)


def test_msm_batch_validation_and_empty_segment(loader):
    s = oracle.synth_scalars(6, 0, 8); p = oracle.synth_points(6, 0, 8, 2)
    with pytest.raises(sv.Error):
        loader.msm_batch(s, p, [0, 3, 3, 8])         # an empty MSM panics in the reference (native.rs:69)
    bad = bytearray(p); bad[64 * 2:64 * 3] = le(1) + le(3)
    with pytest.raises(sv.Error):
        loader.msm_batch(s, bytes(bad), [0, 4, 8], flags=sv.CHECK_INPUTS)


@pytest.mark.parametrize("glv_mode", [1, 2], ids=["glv_always", "glv_never"])
@pytest.mark.parametrize("n", [1, 2, 21, 256, 5000, 70000])
def test_msm_glv_on_off_agree_with_oracle(loader, glv_mode, n):
    """The endomorphism split (csrc/glv.cuh) is an internal choice: forced on and forced off must both reproduce the oracle."""
    s = bytearray(oracle.synth_scalars(81, 0, n)); p = oracle.synth_points(81, 0, n, 8)
    for j, val in enumerate((0, 1, m.R - 1, (m.R - 1) // 2, 0x30644E72E131A029048B6E193FD84104CC37A73FEC2BC5E9B8CA0B2D36636F23)):
        if j < n:
            s[32 * j:32 * j + 32] = le(val)            # 0, 1, r-1, (r-1)/2, lambda itself
    s = bytes(s)
    exp = oracle.msm_pippenger(s, p, n, 8)
    loader.set_glv_mode(glv_mode)
    try:
        assert loader.msm(s, p, n) == exp
    finally:
        loader.set_glv_mode(0)


def test_msm_glv_forced_at_large_size_matches_checksum(loader):
    import torch
    n = 1 << 22
    ds = torch.empty(n * 32, dtype=torch.uint8, device="cuda"); dp = torch.empty(n * 64, dtype=torch.uint8, device="cuda")
    out = torch.zeros(128, dtype=torch.uint8, device="cuda")
    loader.synth_scalars_device(83, 0, n, ds.data_ptr()); loader.synth_points_device(83, 0, n, dp.data_ptr())
    loader.set_glv_mode(1)
    loader.msm_device(ds.data_ptr(), dp.data_ptr(), n, d_out_affine=out.data_ptr())
    loader.set_glv_mode(2)
    loader.msm_device(ds.data_ptr(), dp.data_ptr(), n, d_out_affine=out.data_ptr() + 64)
    loader.set_glv_mode(0)
    torch.cuda.synchronize()
    o = bytes(out.cpu().numpy())
    assert o[:64] == o[64:] == oracle.msm_expected_from_dlogs(ds.cpu().numpy(), oracle.synth_point_scalars(83, 0, n), n)


def test_context_recovers_after_rejected_input(loader):
    """A call that fails validation must leave the context usable (grow-only workspace, stream, status words)."""
    n = 300
    s = oracle.synth_scalars(90, 0, n); p = oracle.synth_points(90, 0, n, 4)
    bad = bytearray(p); bad[64 * 17 + 5] ^= 0x40
    for _ in range(2):
        with pytest.raises(sv.Error):
            loader.msm(s, bytes(bad), n, flags=sv.CHECK_INPUTS)
        assert loader.msm(s, p, n, flags=sv.CHECK_INPUTS) == oracle.msm_pippenger(s, p, n, 4)
    with pytest.raises(sv.Error):
        loader.msm(s, p, 0)
    assert loader.msm(s, p, n) == oracle.msm_pippenger(s, p, n, 4)


def test_instrumentation_api(loader):
    n = 4096
    s = oracle.synth_scalars(91, 0, n); p = oracle.synth_points(91, 0, n, 4)
    before = loader.launch_count
    loader.profile(True)
    try:
        loader.msm(s, p, n)
        stages = loader.stage_times()
    finally:
        loader.profile(False)
    names = [a for a, _, _ in stages]
    assert "msm_bucket_accumulate" in names and "msm_final" in names
    assert all(ms >= 0 for _, ms, _ in stages) and sum(k for _, _, k in stages) == loader.launch_count - before
    plan = loader.msm_plan(n)
    assert plan["buckets_per_window"] == 1 << (plan["window_bits"] - 1) and plan["windows"] * plan["window_bits"] >= 130


# ---------------------------------------------------------------------------------------------------------------------
# SURVEY §8 f1 (first "next" row): Fr scalar preparation on the device
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("which", ["canonical", "montgomery"])
def test_fr_powers_batch_invert_mul_vec(loader, mont_loader, which):
    L = loader if which == "canonical" else mont_loader
    enc = (lambda v: le(v)) if which == "canonical" else (lambda v: le(v * (1 << 256) % m.R))
    rng = np.random.default_rng(12)
    r = int.from_bytes(rng.bytes(31), "little")
    n = 1000
    assert L.powers(enc(r), n) == b"".join(enc(pow(r, i, m.R)) for i in range(n))                 # loader.rs:71-78
    vals = [int.from_bytes(rng.bytes(32), "little") % m.R for _ in range(n)]
    for j in (0, 5, 63, 64, 999):
        vals[j] = 0                                                                                # zeros stay zero (loader.rs:261)
    vals[1], vals[2] = 1, m.R - 1
    got = L.batch_invert(b"".join(map(enc, vals)), n)
    assert got == b"".join(enc(pow(v, -1, m.R) if v else 0) for v in vals)
    coeff = int.from_bytes(rng.bytes(31), "little")
    got = L.batch_invert(b"".join(map(enc, vals)), n, coeff=enc(coeff))                           # util/arithmetic.rs:47-69
    assert got == b"".join(enc(coeff * pow(v, -1, m.R) % m.R if v else 0) for v in vals)
    a = [int.from_bytes(rng.bytes(32), "little") % m.R for _ in range(n)]
    assert L.fr_mul_vec(b"".join(map(enc, a)), b"".join(map(enc, vals)), n) == b"".join(enc(x * y % m.R) for x, y in zip(a, vals))
    assert L.batch_invert(enc(0) * 3, 3) == enc(0) * 3                                             # all-zero input
