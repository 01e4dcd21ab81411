"""CPU: the Poseidon restatement (oracle/poseidon_model.py: Grain-generated constants, permutation, the reference's sponge and native
transcript) against PUBLIC known answers — the Poseidon reference test vectors poseidonperm_x5_254_3 / poseidonperm_x5_254_5 over the
BN254 scalar field, the vectors the un-vendored `poseidon` crate (snark-verifier/Cargo.toml:28) is itself tested against — plus the
generated constant table the device uses (csrc/poseidon_consts.inc)."""
import os
import re

from oracle import poseidon_model as pm

R = pm.R


def test_public_permutation_vectors_pin_grain_and_permutation():
    rc, mds = pm.generate(3, 8, 57)
    assert rc[0][0] == 0x0EE9A592BA9A9518D05986D656F40C2114C4993C11BB29938D21D47304CD8E6E     # first round constant of x5_254_3
    assert pm.permute([0, 1, 2], rc, mds, 8, 57) == [
        0x115CC0F5E7D690413DF64C6B9662E9CF2A3617F2743245519E19607A4417189A,
        0x0FCA49B798923AB0239DE1C9E7A4A9A2210312B6A2F616D18B5A87F9B628AE29,
        0x0E7AE82E40091E63CBD4F16A6D16310B3729D4B6E138FCF54110E2867045A30C]
    rc, mds = pm.generate(5, 8, 60)
    assert pm.permute([0, 1, 2, 3, 4], rc, mds, 8, 60) == [
        0x299C867DB6C1FDD79DCEFA40E4510B9837E60EBB1CE0663DBAA525DF65250465,
        0x1148AAEF609AA338B27DAFD89BB98862D8BB2B429ACEAC47D86206154FFE053D,
        0x24FEBB87FED7462E23F6665FF9A0111F4044C38EE1672C1AC6B0637D34F24907,
        0x0EB08F6D809668A981C186BEAF6110060707059576406B248E5D9CF6E78B3D3E,
        0x07748BC6877C9B82C8B98666EE9D0626EC7F5BE4205F79EE8528EF1C4A376FC7]


def test_sponge_padding_rules():
    """util/hash/poseidon.rs:159-173: a padding 1 after the last input; an extra permutation when the buffer is a multiple of RATE"""
    rc, mds = pm.spec(5, 8, 60)
    h = pm.Poseidon()
    h.update([7, 8])
    st = [1 << 64, 7, 8, 1, 0]
    assert h.squeeze() == pm.permute(st, rc, mds, 8, 60)[1]
    h = pm.Poseidon()
    h.update([1, 2, 3, 4])
    st = pm.permute([1 << 64, 1, 2, 3, 4], rc, mds, 8, 60)
    st[1] = (st[1] + 1) % R
    assert h.squeeze() == pm.permute(st, rc, mds, 8, 60)[1]
    h2 = pm.Poseidon()                                            # squeezing twice in a row: the second absorbs an empty buffer
    a = h2.squeeze()
    st = pm.permute([1 << 64, 1, 0, 0, 0], rc, mds, 8, 60)
    assert a == st[1]
    st[1] = (st[1] + 1) % R
    assert h2.squeeze() == pm.permute(st, rc, mds, 8, 60)[1]


def test_generated_device_constants_match_the_model():
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "snark_verifier_b200", "csrc", "poseidon_consts.inc")
    txt = open(path).read()
    rc, mds = pm.spec(5, 8, 60)
    mont = lambda v: (v << 256) % R
    limbs = lambda v: [(mont(v) >> (32 * i)) & 0xFFFFFFFF for i in range(8)]
    body = txt[txt.index("SNARKV_POSEIDON_RC_INIT"):]
    words = [int(w, 16) for w in re.findall(r"0x([0-9a-f]{8})u", body)]
    want = [l for row in rc for v in row for l in limbs(v)] + [l for row in mds for v in row for l in limbs(v)]
    assert words == want


def test_lane_schedule_of_the_device_permutation_matches_the_plain_rounds():
    """csrc/poseidon.cu poseidon_permute_lanes: one proof per group of five lanes, lane g holds state[g].  A partial round is six
    multiplication steps — lane 0: x^2, x^4, x^5; lanes j >= 1: M[0][j] x_j (an entry of lane 0's row), then M[j][1..4] x_1..4; last step
    every lane M[g][0] x_0^5; lane 0 sums the four row-0 entries it is handed, lanes j their own four.  Emulated lane by lane over
    exact integers (every value a lane reads from another lane goes through `shfl`) against the textbook rounds."""
    T, RF, RP = 5, 8, 60
    rc, mds = pm.generate(T, RF, RP)

    def permute_lanes(state):
        s = list(state)                                   # s[g] = register of lane g
        for r in range(RF + RP):
            x = [(s[g] + rc[r][g]) % R for g in range(T)]
            shfl = lambda v, lane: v[lane]                # noqa: E731 — a value read from another lane of the group
            if r < RF // 2 or r >= RF // 2 + RP:
                y = [pow(x[g], 5, R) for g in range(T)]
                s = [sum(mds[g][j] * shfl(y, j) for j in range(T)) % R for g in range(T)]
                continue
            t1, t2, t3, t4, t5 = [0] * T, [0] * T, [0] * T, [0] * T, [0] * T
            for g in range(T):
                z = g == 0
                t1[g] = (x[g] if z else mds[0][g]) * x[g] % R                                   # step 1
                t2[g] = (t1[g] if z else mds[g][1]) * (t1[g] if z else shfl(x, 1)) % R          # step 2
                t3[g] = (t2[g] if z else mds[g][2]) * (x[g] if z else shfl(x, 2)) % R           # step 3
                t4[g] = mds[g][3] * shfl(x, 3) % R                                              # steps 4, 5 (lane 0 idles along)
                t5[g] = mds[g][4] * shfl(x, 4) % R
            q = [mds[g][0] * shfl(t3, 0) % R for g in range(T)]                                 # step 6
            row0 = (shfl(t1, 1) + shfl(t1, 2) + shfl(t1, 3) + shfl(t1, 4)) % R
            s = [(q[g] + (row0 if g == 0 else t2[g] + t3[g] + t4[g] + t5[g])) % R for g in range(T)]
        return s

    for state in ([0, 1, 2, 3, 4], [R - 1, 0, 12345, 1 << 200, 7]):
        assert permute_lanes(state) == pm.permute(list(state), rc, mds, RF, RP)
