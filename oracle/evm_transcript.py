"""CPU restatement of the reference's `EvmTranscript<_, NativeLoader, _, _>` (TEST INFRASTRUCTURE ONLY).

Reference: /root/reference/snark-verifier/src/system/halo2/transcript/evm.rs:175-222 and loader/evm/util.rs:61-67 (`u256_to_fe`):
  * `common_scalar`   appends the 32-byte BIG-endian repr of the scalar to `buf`                       (:219-222)
  * `common_ec_point` appends x then y, each 32-byte big-endian                                          (:200-216)
  * `squeeze_challenge`: data = buf ++ [0x01 if len(buf) == 0x20]; hash = Keccak256(data); buf = hash;
    challenge = U256::from_be_bytes(hash) % r                                                            (:184-198)

Keccak-256 here is the original Keccak padding (0x01), NOT NIST SHA3-256 (0x06); Python's hashlib only ships the latter, so the
permutation is written out below and pinned against hashlib.sha3_256 (same permutation, other padding byte) plus the well-known
Keccak-256 digests of b"" and b"abc" (tests/test_evm_transcript.py).
"""
MASK = (1 << 64) - 1
R_MOD = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001

RC = [0x0000000000000001, 0x0000000000008082, 0x800000000000808A, 0x8000000080008000, 0x000000000000808B, 0x0000000080000001,
      0x8000000080008081, 0x8000000000008009, 0x000000000000008A, 0x0000000000000088, 0x0000000080008009, 0x000000008000000A,
      0x000000008000808B, 0x800000000000008B, 0x8000000000008089, 0x8000000000008003, 0x8000000000008002, 0x8000000000000080,
      0x000000000000800A, 0x800000008000000A, 0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008]
# rotation offsets r[x][y]
ROT = [[0, 36, 3, 41, 18], [1, 44, 10, 45, 2], [62, 6, 43, 15, 61], [28, 55, 25, 21, 56], [27, 20, 39, 8, 14]]


def _rol(v, n):
    n %= 64
    return ((v << n) | (v >> (64 - n))) & MASK if n else v


def keccak_f1600(a):
    """a: list of 25 lanes, index x + 5 y."""
    for rnd in range(24):
        c = [a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20] for x in range(5)]
        d = [c[(x - 1) % 5] ^ _rol(c[(x + 1) % 5], 1) for x in range(5)]
        a = [a[i] ^ d[i % 5] for i in range(25)]
        b = [0] * 25
        for x in range(5):
            for y in range(5):
                b[y + 5 * ((2 * x + 3 * y) % 5)] = _rol(a[x + 5 * y], ROT[x][y])
        a = [b[i] ^ ((~b[(i % 5 + 1) % 5 + 5 * (i // 5)]) & MASK & b[(i % 5 + 2) % 5 + 5 * (i // 5)]) for i in range(25)]
        a[0] ^= RC[rnd]
    return a


def _sponge256(data, pad_byte):
    rate = 136
    msg = bytearray(data)
    msg.append(pad_byte)
    while len(msg) % rate:
        msg.append(0)
    msg[-1] |= 0x80
    st = [0] * 25
    for off in range(0, len(msg), rate):
        for i in range(rate // 8):
            st[i] ^= int.from_bytes(msg[off + 8 * i:off + 8 * i + 8], "little")
        st = keccak_f1600(st)
    return b"".join(st[i].to_bytes(8, "little") for i in range(4))


def keccak256(data):
    return _sponge256(data, 0x01)


def sha3_256(data):
    return _sponge256(data, 0x06)


class EvmTranscript:
    """Native-loader EvmTranscript: the byte buffer, `common_*` and `squeeze_challenge` exactly as evm.rs:184-222."""

    def __init__(self):
        self.buf = b""

    def common_scalar(self, scalar_int):
        self.buf += int(scalar_int).to_bytes(32, "big")

    def common_ec_point(self, x_int, y_int):
        self.buf += int(x_int).to_bytes(32, "big") + int(y_int).to_bytes(32, "big")

    def common_bytes32(self, word):
        assert len(word) == 32
        self.buf += bytes(word)

    def squeeze_challenge(self):
        data = self.buf + (b"\x01" if len(self.buf) == 0x20 else b"")
        h = keccak256(data)
        self.buf = h
        return int.from_bytes(h, "big") % R_MOD


def challenges_for_stream(stream, seg_end):
    """The challenge sequence for one absorbed byte stream cut at the byte offsets `seg_end` (one squeeze after each)."""
    t = EvmTranscript()
    out, prev = [], 0
    for e in seg_end:
        for off in range(prev, e, 32):
            t.common_bytes32(stream[off:off + 32])
        out.append(t.squeeze_challenge())
        prev = e
    return out
