"""SURVEY §8 f2 (second "next" row): the Keccak-256 `EvmTranscript` (transcript/evm.rs:184-222) — CPU oracle pinned, and the device
batch kernel (one thread per proof) against it."""
import hashlib

import numpy as np
import pytest

from oracle import evm_transcript as et


def test_python_keccak_permutation_is_pinned():
    rng = np.random.default_rng(1)
    for ln in (0, 1, 31, 32, 135, 136, 137, 272, 1000):
        data = rng.bytes(ln)
        assert et.sha3_256(data) == hashlib.sha3_256(data).digest()         # same permutation + sponge, NIST padding
    # original-Keccak padding: the digests every Ethereum tool agrees on
    assert et.keccak256(b"").hex() == "c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470"
    assert et.keccak256(b"abc").hex() == "4e03657aea45a94fc7d47ba826c8d667c0d1e6e33a64a036ec44f58fa12d6c45"


def test_transcript_semantics():
    t = et.EvmTranscript()
    t.common_scalar(5)
    c0 = t.squeeze_challenge()                                              # buf is exactly 32 bytes -> 0x01 appended (evm.rs:190-194)
    assert c0 == int.from_bytes(et.keccak256((5).to_bytes(32, "big") + b"\x01"), "big") % et.R_MOD
    c1 = t.squeeze_challenge()                                              # nothing absorbed in between: buf = previous hash (32 B) -> 0x01 again
    assert c1 == int.from_bytes(et.keccak256(et.keccak256((5).to_bytes(32, "big") + b"\x01") + b"\x01"), "big") % et.R_MOD
    t.common_ec_point(1, 2)
    c2 = t.squeeze_challenge()                                              # 96 bytes: no extra byte
    assert c2 < et.R_MOD and c2 != c1


@pytest.mark.gpu
def test_device_transcript_batch_matches_oracle():
    import snark_verifier_b200 as sv
    rng = np.random.default_rng(2)
    L = sv.CudaLoader(0)
    try:
        for m_proofs, words, seg_words in ((1, 1, [1]), (3, 4, [1, 1, 4]), (257, 40, [1, 9, 9, 12, 30, 30, 40]), (64, 80, [0, 17, 34, 34, 80])):
            stream_len = 32 * words
            streams = rng.bytes(m_proofs * stream_len) if stream_len else b""
            seg_end = [32 * w for w in seg_words]
            got = L.evm_transcript_challenges(streams, stream_len, seg_end, m_proofs)
            k = len(seg_end)
            for j in range(m_proofs):
                exp = et.challenges_for_stream(streams[j * stream_len:(j + 1) * stream_len], seg_end)
                for i in range(k):
                    assert got[(j * k + i) * 32:(j * k + i + 1) * 32] == exp[i].to_bytes(32, "little"), (m_proofs, j, i)
    finally:
        L.close()
