"""GPU parity of csrc/poseidon.cu: the device Poseidon permutation against the PUBLIC test vector poseidonperm_x5_254_5,
batched transcript challenges against the CPU restatement of the reference's sponge / native PoseidonTranscript
(oracle/poseidon_model.py), and compressed-point parsing (`C::from_bytes`) against Python big integers.  Bar: bit-exact."""
import random

import pytest

import oracle
import snark_verifier_b200 as sv
from oracle import bn254_model as m
from oracle import poseidon_model as pos

pytestmark = pytest.mark.gpu
R = m.R
le = m.fe_to_le


@pytest.fixture(scope="module", params=[sv.CANONICAL, sv.MONTGOMERY], ids=["canonical", "montgomery"])
def loader(request):
    L = sv.CudaLoader(0, fmt=request.param)
    yield L
    L.close()


def enc(L, v):
    return le(v if L.fmt == sv.CANONICAL else (v << 256) % R)


def dec(L, b):
    v = int.from_bytes(b, "little")
    return v if L.fmt == sv.CANONICAL else v * pow(1 << 256, -1, R) % R


def test_device_permutation_reproduces_the_public_vector(loader):
    out = loader.poseidon_permute(b"".join(enc(loader, v) for v in range(5)), 1)
    assert [dec(loader, out[32 * i:32 * i + 32]) for i in range(5)] == [
        0x299C867DB6C1FDD79DCEFA40E4510B9837E60EBB1CE0663DBAA525DF65250465,
        0x1148AAEF609AA338B27DAFD89BB98862D8BB2B429ACEAC47D86206154FFE053D,
        0x24FEBB87FED7462E23F6665FF9A0111F4044C38EE1672C1AC6B0637D34F24907,
        0x0EB08F6D809668A981C186BEAF6110060707059576406B248E5D9CF6E78B3D3E,
        0x07748BC6877C9B82C8B98666EE9D0626EC7F5BE4205F79EE8528EF1C4A376FC7]
    rnd = random.Random(4)
    states = [[rnd.randrange(R) for _ in range(5)] for _ in range(70)] + [[0] * 5, [R - 1] * 5]
    out = loader.poseidon_permute(b"".join(enc(loader, v) for s in states for v in s), len(states))
    rc, mds = pos.spec(5, 8, 60)
    for j, s in enumerate(states):
        assert [dec(loader, out[32 * (5 * j + i):32 * (5 * j + i + 1)]) for i in range(5)] == pos.permute(s, rc, mds, 8, 60)


@pytest.mark.parametrize("seg_end", [[0], [1, 1, 1], [4, 8, 9], [3, 3, 10, 17, 17, 30], [5, 13, 14, 14, 33, 40]])
def test_transcript_challenges_match_the_reference_sponge(loader, seg_end):
    """every padding case of Poseidon::squeeze (util/hash/poseidon.rs:159-173): empty buffer, exact multiples of RATE, remainders"""
    rnd = random.Random(sum(seg_end))
    mm, ln = 37, seg_end[-1] + 2                                # two trailing elements no squeeze ever absorbs
    streams = [[rnd.randrange(R) for _ in range(ln)] for _ in range(mm)]
    got = loader.poseidon_transcript_challenges(b"".join(enc(loader, v) for s in streams for v in s), ln, seg_end, mm)
    k = len(seg_end)
    for j, s in enumerate(streams):
        assert [dec(loader, got[32 * (j * k + i):32 * (j * k + i + 1)]) for i in range(k)] == pos.challenges_for_elements(s, seg_end)


def test_native_transcript_order_scalars_and_points(loader):
    """common_scalar / common_ec_point / squeeze_challenge of the native PoseidonTranscript (transcript/halo2.rs:201-242)"""
    tr = pos.PoseidonTranscript()
    rnd = random.Random(9)
    elements, seg_end, want = [], [], []
    for rounds in range(4):
        for _ in range(rnd.randrange(0, 4)):
            v = rnd.randrange(R)
            tr.common_scalar(v); elements.append(v)
        for _ in range(rnd.randrange(0, 3)):
            pt = m.g1_mul(m.G1_GEN, rnd.randrange(1, R))
            tr.common_ec_point(*pt); elements += [pt[0] % R, pt[1] % R]
        seg_end.append(len(elements))
        want.append(tr.squeeze_challenge())
    got = loader.poseidon_transcript_challenges(b"".join(enc(loader, v) for v in elements), len(elements), seg_end, 1)
    assert [dec(loader, got[32 * i:32 * i + 32]) for i in range(4)] == want


def test_compressed_points_from_bytes(loader):
    rnd = random.Random(12)
    P = m.P
    pts = [m.g1_mul(m.G1_GEN, rnd.randrange(1, R)) for _ in range(40)] + [m.G1_GEN, m.g1_neg(m.G1_GEN)]
    comp = [(x | ((y & 1) << 255)).to_bytes(32, "little") for x, y in pts]
    bad = [(1 << 254).to_bytes(32, "little"),                                  # identity flag
           bytes(32),                                                          # all zeros: x = 0 -> y^2 = 3 (no point with the even root? decided by the model below)
           (P).to_bytes(32, "little"),                                         # x = p: not canonical
           (P + 1).to_bytes(32, "little")]
    x = 2
    while pow((x ** 3 + 3) % P, (P - 1) // 2, P) == 1:                         # an x whose x^3 + 3 is a non-residue
        x += 1
    bad.append(x.to_bytes(32, "little"))
    points, els, valid = loader.g1_decompress(b"".join(comp + bad), len(comp) + len(bad))
    fq = lambda v: le(v if loader.fmt == sv.CANONICAL else (v << 256) % P)
    for i, (px, py) in enumerate(pts):
        assert valid[i] == 1
        assert points[64 * i:64 * i + 64] == fq(px) + fq(py)
        assert els[64 * i:64 * i + 64] == enc(loader, px % R) + enc(loader, py % R)
    n = len(pts)
    y0 = pow(3, (P + 1) // 4, P)
    zero_is_point = y0 * y0 % P == 3                                           # (0, sqrt 3) would be a point if 3 were a residue
    assert list(valid[n:]) == [0, 1 if zero_is_point else 0, 0, 0, 0]
    assert points[64 * n:64 * n + 64] == bytes(64)
    # (a coordinate in [r, p) would be absorbed reduced — fe_to_fe — but p - r ~ 2^127: no such point can be found to test with)


def test_huge_batch_uses_the_thread_per_proof_kernel(loader):
    """above 2^16 proofs the transcript runs one proof per thread instead of one per 8 lanes: same challenges"""
    import numpy as np
    if loader.fmt != sv.CANONICAL:
        pytest.skip("one format is enough for the dispatch test")
    mm, ln, seg_end = (1 << 16) + 37, 6, [2, 6]
    rng = np.random.default_rng(3)
    st = rng.integers(0, 256, size=(mm, ln, 32), dtype=np.uint8)
    st[:, :, 31] &= 0x1F                                        # < 2^253 < r
    got = loader.poseidon_transcript_challenges(st.tobytes(), ln, seg_end, mm)
    for j in (0, 1, 4095, 65535, mm - 1):
        el = [int.from_bytes(st[j, i].tobytes(), "little") for i in range(ln)]
        assert [int.from_bytes(got[32 * (2 * j + i):32 * (2 * j + i + 1)], "little") for i in range(2)] == pos.challenges_for_elements(el, seg_end)
