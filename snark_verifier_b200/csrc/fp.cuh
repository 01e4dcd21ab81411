// fp.cuh — BN254 base/scalar field arithmetic for sm_100a, 8 x 32-bit limbs in Montgomery form (R = 2^256).
//
// Replaces (on the device) halo2curves 0.6.0 `bn256::{Fq, Fr}`, the arithmetic that the reference reaches through
// snark-verifier/src/util/arithmetic.rs:13-23 and calls from loader/native.rs:67-70 and pcs/kzg/decider.rs:74-78.
// In-memory layout is identical to halo2curves' `[u64; 4]` little-endian Montgomery limbs, so a `&[G1Affine]` /
// `&[Fr]` slice from the Rust side can be handed to the kernels byte-for-byte (SNARKV_MONTGOMERY format).
//
// The multiply/add/sub bodies are single inline-PTX blocks generated and CPU-verified by gen_field_ptx.py: two
// interleaved 64-bit-column accumulators so that ptxas fuses every mad.lo.cc/madc.hi.cc pair into one
// IMAD.WIDE.U32.X (≈170 IMAD-pipe issues per Montgomery multiplication; checked with cuobjdump -sass).
#pragma once
#include <cstdint>

#include "fp_inv.cuh"
// Curve configuration of this translation unit.  The default build is BN254 (FQ = base field, FR = scalar field).  A unit that
// defines SNARKV_CURVE_PALLAS before any include (msm_pasta.cu, which also renames the namespace) gets the same code over the
// Pallas fields — SURVEY §8 f4: the IPA decider's MSM (pcs/ipa/decider.rs:47-55) — with y^2 = x^3 + 5 and 255-bit moduli.
#ifdef SNARKV_CURVE_PALLAS
#include "fp_ptx_pallas.inc"
#define SNARKV_FIELD_BITS 255
#define SNARKV_CURVE_B 5
#else
#include "fp_ptx.inc"
#define SNARKV_FIELD_BITS 254
#define SNARKV_CURVE_B 3
#endif

namespace snarkv {

enum Field : int { FQ = 0, FR = 1 };

#define SNARKV_FP_OPS(r, a, b)                                                                                   \
    : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]),          \
      "=r"(r.v[7])                                                                                               \
    : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),    \
      "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7])

#define SNARKV_FP_OPS1(r, a)                                                                                     \
    : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]),          \
      "=r"(r.v[7])                                                                                               \
    : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7])

template <Field F>
struct alignas(16) Fp {
    uint32_t v[8];
};
typedef Fp<FQ> Fq;
typedef Fp<FR> Fr;

// ---- constants (little-endian 32-bit limbs) -------------------------------------------------------------------------
template <Field F> __host__ __device__ __forceinline__ constexpr uint32_t fp_mod_limb(int i);
template <> __host__ __device__ __forceinline__ constexpr uint32_t fp_mod_limb<FQ>(int i) {
    constexpr uint32_t m[8] = SNARKV_FQ_MOD_LIMBS;
    return m[i];
}
template <> __host__ __device__ __forceinline__ constexpr uint32_t fp_mod_limb<FR>(int i) {
    constexpr uint32_t m[8] = SNARKV_FR_MOD_LIMBS;
    return m[i];
}
// R = 2^256 mod m  (Montgomery one)
template <Field F> __host__ __device__ __forceinline__ constexpr uint32_t fp_one_limb(int i);
template <> __host__ __device__ __forceinline__ constexpr uint32_t fp_one_limb<FQ>(int i) {
    constexpr uint32_t m[8] = SNARKV_FQ_ONE_LIMBS;
    return m[i];
}
template <> __host__ __device__ __forceinline__ constexpr uint32_t fp_one_limb<FR>(int i) {
    constexpr uint32_t m[8] = SNARKV_FR_ONE_LIMBS;
    return m[i];
}
// R^2 mod m
template <Field F> __host__ __device__ __forceinline__ constexpr uint32_t fp_r2_limb(int i);
template <> __host__ __device__ __forceinline__ constexpr uint32_t fp_r2_limb<FQ>(int i) {
    constexpr uint32_t m[8] = SNARKV_FQ_R2_LIMBS;
    return m[i];
}
template <> __host__ __device__ __forceinline__ constexpr uint32_t fp_r2_limb<FR>(int i) {
    constexpr uint32_t m[8] = SNARKV_FR_R2_LIMBS;
    return m[i];
}

template <Field F> __device__ __forceinline__ Fp<F> fp_zero() {
    Fp<F> r;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.v[i] = 0;
    return r;
}
template <Field F> __device__ __forceinline__ Fp<F> fp_one() {
    Fp<F> r;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.v[i] = fp_one_limb<F>(i);
    return r;
}
template <Field F> __device__ __forceinline__ Fp<F> fp_r2() {
    Fp<F> r;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.v[i] = fp_r2_limb<F>(i);
    return r;
}

// ---- core ops -----------------------------------------------------------------------------------------------------------
template <Field F> __device__ __forceinline__ Fp<F> fp_mul(const Fp<F>& a, const Fp<F>& b) {
    Fp<F> r;
    if constexpr (F == FQ) asm(SNARKV_PTX_FQ_MUL SNARKV_FP_OPS(r, a, b));
    else asm(SNARKV_PTX_FR_MUL SNARKV_FP_OPS(r, a, b));
    return r;
}
// a b + c d with ONE Montgomery reduction (generated MUL2ADD block: 128 + 64 instead of 256 partial products); inputs fully reduced.
// BN254 only (the row sums need a modulus below 2^254): other curves take the two-multiplication path.
#define SNARKV_FP_OPS4(r, a, b, c, d)                                                                             \
    : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]),          \
      "=r"(r.v[7])                                                                                               \
    : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),    \
      "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]),    \
      "r"(c.v[0]), "r"(c.v[1]), "r"(c.v[2]), "r"(c.v[3]), "r"(c.v[4]), "r"(c.v[5]), "r"(c.v[6]), "r"(c.v[7]),    \
      "r"(d.v[0]), "r"(d.v[1]), "r"(d.v[2]), "r"(d.v[3]), "r"(d.v[4]), "r"(d.v[5]), "r"(d.v[6]), "r"(d.v[7])
template <Field F> __device__ __forceinline__ Fp<F> fp_mul2add(const Fp<F>& a, const Fp<F>& b, const Fp<F>& c, const Fp<F>& d) {
#if defined(SNARKV_PTX_FQ_MUL2ADD) && !defined(SNARKV_NO_MUL2ADD)
    Fp<F> r;
    if constexpr (F == FQ) asm(SNARKV_PTX_FQ_MUL2ADD SNARKV_FP_OPS4(r, a, b, c, d));
    else asm(SNARKV_PTX_FR_MUL2ADD SNARKV_FP_OPS4(r, a, b, c, d));
    return r;
#else
    return fp_add(fp_mul(a, b), fp_mul(c, d));
#endif
}
// a^2: dedicated generated block (gen_mont_sqr: every unordered pair of limbs multiplied once, 36 + 64 instead of 64 + 64 products)
template <Field F> __device__ __forceinline__ Fp<F> fp_sqr(const Fp<F>& a) {
#ifdef SNARKV_NO_DEDICATED_SQR
    return fp_mul(a, a);
#else
    Fp<F> r;
    if constexpr (F == FQ) asm(SNARKV_PTX_FQ_SQR SNARKV_FP_OPS1(r, a));
    else asm(SNARKV_PTX_FR_SQR SNARKV_FP_OPS1(r, a));
    return r;
#endif
}
template <Field F> __device__ __forceinline__ Fp<F> fp_add(const Fp<F>& a, const Fp<F>& b) {
    Fp<F> r;
    if constexpr (F == FQ) asm(SNARKV_PTX_FQ_ADD SNARKV_FP_OPS(r, a, b));
    else asm(SNARKV_PTX_FR_ADD SNARKV_FP_OPS(r, a, b));
    return r;
}
template <Field F> __device__ __forceinline__ Fp<F> fp_sub(const Fp<F>& a, const Fp<F>& b) {
    Fp<F> r;
    if constexpr (F == FQ) asm(SNARKV_PTX_FQ_SUB SNARKV_FP_OPS(r, a, b));
    else asm(SNARKV_PTX_FR_SUB SNARKV_FP_OPS(r, a, b));
    return r;
}
template <Field F> __device__ __forceinline__ Fp<F> fp_dbl(const Fp<F>& a) { return fp_add(a, a); }
template <Field F> __device__ __forceinline__ Fp<F> fp_neg(const Fp<F>& a) { return fp_sub(fp_zero<F>(), a); }
template <Field F> __device__ __forceinline__ bool fp_is_zero(const Fp<F>& a) {
    return (a.v[0] | a.v[1] | a.v[2] | a.v[3] | a.v[4] | a.v[5] | a.v[6] | a.v[7]) == 0;
}
template <Field F> __device__ __forceinline__ bool fp_eq(const Fp<F>& a, const Fp<F>& b) {
    uint32_t d = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) d |= a.v[i] ^ b.v[i];
    return d == 0;
}
// canonical integer (< m) -> Montgomery form, and back
template <Field F> __device__ __forceinline__ Fp<F> fp_to_mont(const Fp<F>& a) { return fp_mul(a, fp_r2<F>()); }
// a / 2^256 mod m: the Montgomery reduction alone (generated FROM_MONT block: the products by the constant 1 are moves / carry
// chains, half the multiplier work of fp_mul(a, 1) — k_digits converts every scalar of a Montgomery-format MSM with it)
template <Field F> __device__ __forceinline__ Fp<F> fp_from_mont(const Fp<F>& a) {
    Fp<F> r;
    if constexpr (F == FQ) asm(SNARKV_PTX_FQ_FROM_MONT SNARKV_FP_OPS1(r, a));
    else asm(SNARKV_PTX_FR_FROM_MONT SNARKV_FP_OPS1(r, a));
    return r;
}
// true iff the raw 256-bit value is a canonical residue (< m) — `PrimeField::from_repr` acceptance test
template <Field F> __device__ __forceinline__ bool fp_is_canonical(const Fp<F>& a) {
#pragma unroll
    for (int i = 7; i >= 0; --i) {
        if (a.v[i] < fp_mod_limb<F>(i)) return true;
        if (a.v[i] > fp_mod_limb<F>(i)) return false;
    }
    return false;
}

// a^(m-2) by square-and-multiply over the constant exponent (used once per result for to_affine; not on hot loops)
template <Field F> __device__ __noinline__ Fp<F> fp_inv(const Fp<F>& a) {
    Fp<F> r = fp_one<F>();
    // exponent m - 2, MSB first.  m is odd and m-2 only changes limb 0 (no borrow: low limbs are ...47 / ...01 -> need care)
    uint32_t e[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) e[i] = fp_mod_limb<F>(i);
    // subtract 2 with borrow
    uint32_t borrow = 2;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        uint32_t t = e[i] - borrow;
        borrow = (e[i] < borrow) ? 1u : 0u;
        e[i] = t;
    }
    for (int i = SNARKV_FIELD_BITS - 1; i >= 0; --i) {   // BN254: 254-bit moduli, Pasta: 255-bit
        r = fp_sqr(r);
        if ((e[i >> 5] >> (i & 31)) & 1u) r = fp_mul(r, a);
    }
    return r;
}

// R^3 mod m
template <Field F> __device__ __forceinline__ Fp<F> fp_r3() {
    Fp<F> r;
    if constexpr (F == FQ) { constexpr uint32_t m[8] = SNARKV_FQ_R3_LIMBS;
#pragma unroll
        for (int i = 0; i < 8; ++i) r.v[i] = m[i]; }
    else { constexpr uint32_t m[8] = SNARKV_FR_R3_LIMBS;
#pragma unroll
        for (int i = 0; i < 8; ++i) r.v[i] = m[i]; }
    return r;
}
// Inverse on a serial critical path: binary extended GCD (fp_inv.cuh) on the raw Montgomery value aR gives (aR)^-1;
// one Montgomery multiplication by R^3 turns it into a^-1 R.  Data-dependent control flow: use from ONE thread (or lanes
// holding identical data), not in throughput kernels.
template <Field F> __device__ __noinline__ Fp<F> fp_inv_serial(const Fp<F>& a) {
    U256 x, p;
#pragma unroll
    for (int i = 0; i < 8; ++i) { x.v[i] = a.v[i]; p.v[i] = fp_mod_limb<F>(i); }
    U256 y = u256_inv_mod_fast(x, p);   // fp_inv.cuh: 31 binary-GCD steps at a time (the bit-at-a-time u256_inv_mod gives the same value)
    Fp<F> t;
#pragma unroll
    for (int i = 0; i < 8; ++i) t.v[i] = y.v[i];
    return fp_mul(t, fp_r3<F>());
}

// ---- 128-bit vectorised global memory access ----------------------------------------------------------------------------
template <Field F> __device__ __forceinline__ Fp<F> fp_load(const void* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 lo = __ldg(q), hi = __ldg(q + 1);
    Fp<F> r;
    r.v[0] = lo.x; r.v[1] = lo.y; r.v[2] = lo.z; r.v[3] = lo.w;
    r.v[4] = hi.x; r.v[5] = hi.y; r.v[6] = hi.z; r.v[7] = hi.w;
    return r;
}
// coherent variant: for buffers the running kernel also WRITES (the read-only / ld.global.nc path of fp_load is only legal for
// data that stays constant for the whole kernel)
template <Field F> __device__ __forceinline__ Fp<F> fp_load_rw(const void* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 lo = q[0], hi = q[1];
    Fp<F> r;
    r.v[0] = lo.x; r.v[1] = lo.y; r.v[2] = lo.z; r.v[3] = lo.w;
    r.v[4] = hi.x; r.v[5] = hi.y; r.v[6] = hi.z; r.v[7] = hi.w;
    return r;
}
template <Field F> __device__ __forceinline__ void fp_store(void* p, const Fp<F>& a) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(a.v[0], a.v[1], a.v[2], a.v[3]);
    q[1] = make_uint4(a.v[4], a.v[5], a.v[6], a.v[7]);
}

}  // namespace snarkv
