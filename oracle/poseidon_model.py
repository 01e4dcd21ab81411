"""CPU restatement of the Poseidon hasher and the native `PoseidonTranscript` of the reference (TEST INFRASTRUCTURE ONLY).

Reference (paths relative to /root/reference/snark-verifier/src):
  util/hash/poseidon.rs:117-203   Poseidon<F, L, T, RATE>: `update` buffers, `squeeze` absorbs the buffer RATE elements at a time
                                  (one permutation per chunk, a padding 1 after the last input, one extra permutation when the
                                  buffer length is a multiple of RATE) and returns state[1]
  util/hash/poseidon.rs:46-80     absorb_with_pre_constants: inputs are ADDED to state[1..], the padding 1 to the next position
  system/halo2/transcript/halo2.rs:201-274   native PoseidonTranscript: common_scalar = 1 element, common_ec_point = (x mod r, y mod r),
                                  read_scalar = 32-byte little-endian repr, read_ec_point = 32-byte compressed point
The round constants and the MDS matrix come from the un-vendored crate `poseidon` (tag v2024_01_31, snark-verifier/Cargo.toml:28):
`Spec::new(R_F, R_P)` runs the Grain LFSR of the Poseidon reference implementation.  Restated here from the published algorithm
(Poseidon paper, appendix F / `generate_parameters_grain.sage`): 80-bit state from (field type, s-box, field bits, t, R_F, R_P),
160 warm-up bits, self-shrinking output, most-significant-bit-first integers, rejection sampling for the round constants, NO
rejection (reduction mod r) for the 2t Cauchy values x_0..x_{t-1}, y_0..y_{t-1}, M[i][j] = 1 / (x_i + y_j).  The crate's optimised
form (pre-sparse / sparse matrices, util/hash/poseidon.rs:175-203) computes the same permutation.

PINNED externally: the permutation reproduces the public Poseidon test vectors `poseidonperm_x5_254_3` (t = 3, R_F = 8, R_P = 57)
and `poseidonperm_x5_254_5` (t = 5, R_F = 8, R_P = 60) over the BN254 scalar field — the vectors the `poseidon` crate itself is
tested against (tests/test_poseidon.py)."""
R = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
FIELD_BITS = 254


class Grain:
    def __init__(self, field_bits, t, r_f, r_p):
        bits = []

        def app(n, v):
            for i in range(n - 1, -1, -1):
                bits.append((v >> i) & 1)
        app(2, 1); app(4, 0); app(12, field_bits); app(12, t); app(10, r_f); app(10, r_p); app(30, (1 << 30) - 1)
        assert len(bits) == 80
        self.s = bits
        for _ in range(160):
            self._new_bit()

    def _new_bit(self):
        s = self.s
        b = s[62] ^ s[51] ^ s[38] ^ s[23] ^ s[13] ^ s[0]
        s.pop(0)
        s.append(b)
        return b

    def next_bit(self):
        while True:                                   # self-shrinking: a bit is kept iff the one before it is 1
            b1, b2 = self._new_bit(), self._new_bit()
            if b1:
                return b2

    def next_int(self, nbits):
        v = 0
        for _ in range(nbits):
            v = (v << 1) | self.next_bit()
        return v


def generate(t, r_f, r_p, field_bits=FIELD_BITS, modulus=R):
    """-> (round constants [(r_f + r_p) x t], mds [t x t])"""
    g = Grain(field_bits, t, r_f, r_p)
    rc = []
    for _ in range(r_f + r_p):
        row = []
        for _ in range(t):
            while True:
                v = g.next_int(field_bits)
                if v < modulus:
                    break
            row.append(v)
        rc.append(row)
    vals = [g.next_int(field_bits) % modulus for _ in range(2 * t)]
    assert len(set(vals)) == 2 * t
    xs, ys = vals[:t], vals[t:]
    mds = [[pow((xs[i] + ys[j]) % modulus, -1, modulus) for j in range(t)] for i in range(t)]
    return rc, mds


def permute(state, rc, mds, r_f, r_p, modulus=R):
    t = len(state)
    s = list(state)
    for r in range(r_f + r_p):
        s = [(s[i] + rc[r][i]) % modulus for i in range(t)]
        if r < r_f // 2 or r >= r_f // 2 + r_p:
            s = [pow(x, 5, modulus) for x in s]
        else:
            s[0] = pow(s[0], 5, modulus)
        s = [sum(mds[i][j] * s[j] for j in range(t)) % modulus for i in range(t)]
    return s


_SPEC_CACHE = {}


def spec(t, r_f, r_p):
    key = (t, r_f, r_p)
    if key not in _SPEC_CACHE:
        _SPEC_CACHE[key] = generate(t, r_f, r_p)
    return _SPEC_CACHE[key]


class Poseidon:
    """util/hash/poseidon.rs:117-203 over Python ints (T = RATE + 1; the SDK uses T = 5, RATE = 4, R_F = 8, R_P = 60)."""

    def __init__(self, t=5, rate=4, r_f=8, r_p=60):
        assert t == rate + 1
        self.t, self.rate, self.r_f, self.r_p = t, rate, r_f, r_p
        self.rc, self.mds = spec(t, r_f, r_p)
        self.state = [1 << 64] + [0] * (t - 1)          # poseidon::State::default(): capacity word 2^64
        self.buf = []

    def update(self, elements):
        self.buf.extend(int(e) % R for e in elements)

    def _permutation(self, inputs):
        s = self.state
        for i, v in enumerate(inputs):
            s[1 + i] = (s[1 + i] + v) % R
        if len(inputs) < self.rate:
            s[1 + len(inputs)] = (s[1 + len(inputs)] + 1) % R
        self.state = permute(s, self.rc, self.mds, self.r_f, self.r_p)

    def squeeze(self):
        buf, self.buf = self.buf, []
        exact = len(buf) % self.rate == 0
        for i in range(0, len(buf), self.rate):
            self._permutation(buf[i:i + self.rate])
        if exact:
            self._permutation([])
        return self.state[1]


class PoseidonTranscript:
    """Native PoseidonTranscript (system/halo2/transcript/halo2.rs:201-274): what is absorbed, not how it is serialised."""

    def __init__(self, t=5, rate=4, r_f=8, r_p=60):
        self.h = Poseidon(t, rate, r_f, r_p)

    def common_scalar(self, v):
        self.h.update([v])

    def common_ec_point(self, x, y):
        self.h.update([x % R, y % R])                   # fe_to_fe: base-field coordinates reduced into the scalar field

    def squeeze_challenge(self):
        return self.h.squeeze()


def challenges_for_elements(elements, seg_end, **kw):
    """The challenge sequence for one absorbed ELEMENT stream cut at the element offsets `seg_end` (one squeeze after each)."""
    tr = PoseidonTranscript(**kw)
    out, prev = [], 0
    for e in seg_end:
        tr.h.update(elements[prev:e])
        out.append(tr.squeeze_challenge())
        prev = e
    return out
