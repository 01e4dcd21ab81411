// oracle/bn254.hpp — CPU restatement of the BN254 arithmetic under snark-verifier's native hot path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing in the product path (snark_verifier_b200/, include/) may include, link or
// call this.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.
//
// PARITY STATUS: "parity unpinned" by the reference — the arithmetic lives in the un-vendored crate
// halo2curves 0.6.0 (reference snark-verifier/Cargo.toml:14, re-exported at src/util/arithmetic.rs:13-23) and the
// reference ships no known-answer vectors.  What pins this file instead:
//   * oracle/bn254_model.py — an independent Python big-int model (affine formulas, binary Miller loop, naive
//     f^((p^12-1)/r)); tests/golden/*.json are generated from it and asserted against this oracle, and
//   * algebraic properties (bilinearity, [r]P = O, MSM linearity) in tests/.
// Every output is canonical mathematics (affine G1; GT after the full final exponentiation), so agreement of two
// independent implementations is a bit-exact pin.
//
// Representation follows halo2curves' published layout: Fq/Fr = 4 x u64 little-endian limbs in Montgomery form
// (R = 2^256), G1 Jacobian with a = 0, b = 3, identity affine = (0,0); tower Fq2 = Fq[i]/(i^2+1),
// Fq6 = Fq2[v]/(v^3 - (9+i)), Fq12 = Fq6[w]/(w^2 - v).
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

namespace oracle {

typedef uint64_t u64;
typedef unsigned __int128 u128;

extern thread_local u64 g_mulmod_count;  // instrumented Fq/Fr Montgomery multiplications

struct Modulus {
    u64 m[4];
    u64 inv;     // -m^-1 mod 2^64
    u64 r[4];    // 2^256 mod m      (Montgomery one)
    u64 r2[4];   // 2^512 mod m
};
extern const Modulus FQ_MOD, FR_MOD;

// ---- 256-bit helpers -------------------------------------------------------------------------------------------
static inline bool geq(const u64 a[4], const u64 b[4]) {
    for (int i = 3; i >= 0; --i) {
        if (a[i] > b[i]) return true;
        if (a[i] < b[i]) return false;
    }
    return true;
}
static inline u64 add4(u64 r[4], const u64 a[4], const u64 b[4]) {
    u128 c = 0;
    for (int i = 0; i < 4; ++i) { c += (u128)a[i] + b[i]; r[i] = (u64)c; c >>= 64; }
    return (u64)c;
}
static inline u64 sub4(u64 r[4], const u64 a[4], const u64 b[4]) {
    u64 borrow = 0;
    for (int i = 0; i < 4; ++i) {
        u128 d = (u128)a[i] - b[i] - borrow;
        r[i] = (u64)d;
        borrow = (u64)(d >> 64) & 1;
    }
    return borrow;
}

// ---- prime field element in Montgomery form ------------------------------------------------------------------------
template <const Modulus& M>
struct Fp {
    u64 v[4];

    static Fp zero() { Fp r; memset(r.v, 0, 32); return r; }
    static Fp one() { Fp r; memcpy(r.v, M.r, 32); return r; }
    static Fp from_raw(const u64 x[4]) {  // canonical integer (must be < m) -> Montgomery
        Fp a, b; memcpy(a.v, x, 32); memcpy(b.v, M.r2, 32);
        return a * b;
    }
    static Fp from_u64(u64 x) { u64 t[4] = {x, 0, 0, 0}; return from_raw(t); }
    // 32 canonical little-endian bytes (PrimeField::from_repr); returns false if >= m
    static bool from_le_bytes(const uint8_t* b, Fp& out) {
        u64 t[4]; memcpy(t, b, 32);
        if (geq(t, M.m)) return false;
        out = from_raw(t);
        return true;
    }
    void to_raw(u64 out[4]) const {  // Montgomery -> canonical integer (PrimeField::to_repr)
        Fp one_raw; memset(one_raw.v, 0, 32); one_raw.v[0] = 1;
        Fp t = (*this) * one_raw;
        memcpy(out, t.v, 32);
    }
    void to_le_bytes(uint8_t* b) const { u64 t[4]; to_raw(t); memcpy(b, t, 32); }

    bool is_zero() const { return (v[0] | v[1] | v[2] | v[3]) == 0; }
    bool operator==(const Fp& o) const { return memcmp(v, o.v, 32) == 0; }
    bool operator!=(const Fp& o) const { return !(*this == o); }

    Fp operator+(const Fp& o) const {
        Fp r; add4(r.v, v, o.v);           // m < 2^254 so no carry out
        if (geq(r.v, M.m)) sub4(r.v, r.v, M.m);
        return r;
    }
    Fp operator-(const Fp& o) const {
        Fp r;
        if (sub4(r.v, v, o.v)) add4(r.v, r.v, M.m);
        return r;
    }
    Fp neg() const { return is_zero() ? *this : Fp{{0, 0, 0, 0}} - *this; }
    Fp dbl() const { return *this + *this; }
    // CIOS Montgomery multiplication, 4 x 64-bit limbs
    Fp operator*(const Fp& o) const {
        ++g_mulmod_count;
        u64 t[6] = {0, 0, 0, 0, 0, 0};
        for (int i = 0; i < 4; ++i) {
            u128 c = 0;
            for (int j = 0; j < 4; ++j) {
                c += (u128)v[j] * o.v[i] + t[j];
                t[j] = (u64)c; c >>= 64;
            }
            c += t[4]; t[4] = (u64)c; t[5] = (u64)(c >> 64);
            u64 q = t[0] * M.inv;
            c = (u128)q * M.m[0] + t[0]; c >>= 64;
            for (int j = 1; j < 4; ++j) {
                c += (u128)q * M.m[j] + t[j];
                t[j - 1] = (u64)c; c >>= 64;
            }
            c += t[4]; t[3] = (u64)c; t[4] = t[5] + (u64)(c >> 64);
        }
        Fp r; memcpy(r.v, t, 32);
        if (t[4] || geq(r.v, M.m)) sub4(r.v, r.v, M.m);
        return r;
    }
    Fp sqr() const { return (*this) * (*this); }
    Fp pow(const u64 e[4]) const {
        Fp r = one();
        for (int i = 255; i >= 0; --i) {
            r = r.sqr();
            if ((e[i / 64] >> (i % 64)) & 1) r = r * (*this);
        }
        return r;
    }
    Fp inv() const {  // Fermat; inverse of zero is zero
        u64 e[4]; u64 two[4] = {2, 0, 0, 0};
        sub4(e, M.m, two);
        return pow(e);
    }
};

typedef Fp<FQ_MOD> Fq;
typedef Fp<FR_MOD> Fr;

// ---- G1: y^2 = x^3 + 3 -----------------------------------------------------------------------------------------------
struct G1Affine {
    Fq x, y;  // identity = (0, 0)
    bool is_identity() const { return x.is_zero() && y.is_zero(); }
    static G1Affine identity() { return {Fq::zero(), Fq::zero()}; }
    static G1Affine generator() { return {Fq::from_u64(1), Fq::from_u64(2)}; }
    bool is_on_curve() const;
    bool operator==(const G1Affine& o) const { return x == o.x && y == o.y; }
    G1Affine neg() const { return {x, y.neg()}; }
};
struct G1 {  // Jacobian (X/Z^2, Y/Z^3); identity <=> Z == 0
    Fq x, y, z;
    static G1 identity() { return {Fq::zero(), Fq::one(), Fq::zero()}; }
    static G1 from_affine(const G1Affine& p) {
        if (p.is_identity()) return identity();
        return {p.x, p.y, Fq::one()};
    }
    bool is_identity() const { return z.is_zero(); }
    G1 dbl() const;                       // dbl-2009-l
    G1 add(const G1& o) const;            // add-2007-bl
    G1 add_mixed(const G1Affine& o) const;  // madd-2007-bl
    G1Affine to_affine() const;
};
// `impl Mul<Fr> for G1Affine` as halo2curves publishes it: MSB-first over the canonical repr, top bit skipped,
// one double and one (always computed, conditionally selected) mixed add per bit — the cost model of
// NativeLoader::multi_scalar_multiplication (reference loader/native.rs:67).
G1 g1_mul_ct(const G1Affine& p, const Fr& s);
// variable-time double-and-add on a raw 256-bit little-endian integer (test-data generation)
G1 g1_mul_vartime(const G1Affine& p, const u64 k[4]);
void g1_batch_to_affine(const G1* in, G1Affine* out, size_t n);

// reference loader/native.rs:61-71 — fold of base*scalar, one to_affine.  n must be > 0.
G1Affine msm_native_fold(const Fr* scalars, const G1Affine* bases, size_t n);
// reference util/msm.rs:259-304 — serial Pippenger accumulating into *result
void msm_pippenger_serial(const Fr* scalars, const G1Affine* bases, size_t n, G1* result);
// reference util/msm.rs:308-343 — `parallel` feature: chunk per thread + fold
G1 msm_pippenger(const Fr* scalars, const G1Affine* bases, size_t n, int threads);

// ---- tower ----------------------------------------------------------------------------------------------------------
struct Fq2 {
    Fq c0, c1;
    static Fq2 zero() { return {Fq::zero(), Fq::zero()}; }
    static Fq2 one() { return {Fq::one(), Fq::zero()}; }
    bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
    bool operator==(const Fq2& o) const { return c0 == o.c0 && c1 == o.c1; }
    Fq2 operator+(const Fq2& o) const { return {c0 + o.c0, c1 + o.c1}; }
    Fq2 operator-(const Fq2& o) const { return {c0 - o.c0, c1 - o.c1}; }
    Fq2 neg() const { return {c0.neg(), c1.neg()}; }
    Fq2 dbl() const { return {c0.dbl(), c1.dbl()}; }
    Fq2 conj() const { return {c0, c1.neg()}; }
    Fq2 operator*(const Fq2& o) const {  // Karatsuba, 3 Fq muls
        Fq a = c0 * o.c0, b = c1 * o.c1;
        Fq c = (c0 + c1) * (o.c0 + o.c1);
        return {a - b, c - a - b};
    }
    Fq2 sqr() const {  // (c0+c1)(c0-c1), 2 c0 c1
        Fq t = c0 * c1;
        return {(c0 + c1) * (c0 - c1), t.dbl()};
    }
    Fq2 scale(const Fq& k) const { return {c0 * k, c1 * k}; }
    Fq2 mul_xi() const {  // * (9 + i)
        Fq t0 = c0.dbl().dbl().dbl() + c0;  // 9 c0
        Fq t1 = c1.dbl().dbl().dbl() + c1;  // 9 c1
        return {t0 - c1, t1 + c0};
    }
    Fq2 inv() const {
        Fq n = (c0.sqr() + c1.sqr()).inv();
        return {c0 * n, (c1 * n).neg()};
    }
};
struct Fq6 {
    Fq2 c0, c1, c2;
    static Fq6 zero() { return {Fq2::zero(), Fq2::zero(), Fq2::zero()}; }
    static Fq6 one() { return {Fq2::one(), Fq2::zero(), Fq2::zero()}; }
    bool operator==(const Fq6& o) const { return c0 == o.c0 && c1 == o.c1 && c2 == o.c2; }
    Fq6 operator+(const Fq6& o) const { return {c0 + o.c0, c1 + o.c1, c2 + o.c2}; }
    Fq6 operator-(const Fq6& o) const { return {c0 - o.c0, c1 - o.c1, c2 - o.c2}; }
    Fq6 neg() const { return {c0.neg(), c1.neg(), c2.neg()}; }
    Fq6 mul_v() const { return {c2.mul_xi(), c0, c1}; }
    Fq6 operator*(const Fq6& o) const;
    Fq6 sqr() const { return (*this) * (*this); }
    Fq6 mul_by_01(const Fq2& b0, const Fq2& b1) const;
    Fq6 scale(const Fq2& k) const { return {c0 * k, c1 * k, c2 * k}; }
    Fq6 inv() const;
};
struct Fq12 {
    Fq6 c0, c1;
    static Fq12 one() { return {Fq6::one(), Fq6::zero()}; }
    bool operator==(const Fq12& o) const { return c0 == o.c0 && c1 == o.c1; }
    bool is_one() const { return *this == one(); }
    Fq12 operator*(const Fq12& o) const;
    Fq12 sqr() const;
    Fq12 conj() const { return {c0, c1.neg()}; }
    Fq12 inv() const;
    Fq12 frobenius(int power) const;  // power in 1..3
    Fq12 cyclotomic_sqr() const;      // valid only in the cyclotomic subgroup (after the easy part)
    // sparse multiply by  l0 + l3 w + l4 w^3   (tower slots c0.c0, c1.c0, c1.c1)
    Fq12 mul_by_034(const Fq2& l0, const Fq2& l3, const Fq2& l4) const;
    void to_le_bytes(uint8_t out[384]) const;
};

// ---- G2 + pairing -------------------------------------------------------------------------------------------------------
struct G2Affine {
    Fq2 x, y;  // identity = (0,0)
    bool is_identity() const { return x.is_zero() && y.is_zero(); }
    static G2Affine generator();
    bool is_on_curve() const;
    G2Affine neg() const { return {x, y.neg()}; }
};
G2Affine g2_mul_vartime(const G2Affine& q, const u64 k[4]);

struct LineCoeff { Fq2 cy, cx, c0; };  // l(P) = cy*yP + cx*xP * w + c0 * w^3
// `G2Prepared::from(G2Affine)`: all line coefficients of the optimal-ate loop for a fixed Q
struct G2Prepared {
    std::vector<LineCoeff> coeffs;
    bool infinity;
    explicit G2Prepared(const G2Affine& q);
};
// `Bn256::multi_miller_loop(&[(&G1Affine, &G2Prepared)])` — shared squarings, identity pairs skipped
Fq12 multi_miller_loop(const G1Affine* const* ps, const G2Prepared* const* qs, size_t n);
// `MillerLoopResult::final_exponentiation` — f^((p^12-1)/r)
Fq12 final_exponentiation(const Fq12& f);

// reference pcs/kzg/decider.rs:70-82 — G2Prepared::from for g2 and -s_g2 is recomputed on every call, as there.
bool kzg_decide(const G1Affine& lhs, const G1Affine& rhs, const G2Affine& g2, const G2Affine& s_g2, Fq12* gt_out);

void init();  // derive Frobenius constants etc.; idempotent, called by every C entry point

}  // namespace oracle
