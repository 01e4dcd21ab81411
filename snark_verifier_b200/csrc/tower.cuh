// tower.cuh — Fq2 / Fq6 / Fq12 arithmetic for the BN254 pairing kernels (sm_100a).
//
// Device-side replacement for halo2curves 0.6.0 `bn256::{Fq2, Fq6, Fq12}` reached from the reference at
// snark-verifier/src/pcs/kzg/decider.rs:74-78 (`multi_miller_loop`, `final_exponentiation`, `is_identity`).
// Tower: Fq2 = Fq[i]/(i^2+1), Fq6 = Fq2[v]/(v^3 - xi), Fq12 = Fq6[w]/(w^2 - v), xi = 9 + i.
//
// Code-size policy: one thread runs a whole pairing check, so the Fq2 multiply/square are kept as real calls
// (__noinline__, ~3 Montgomery multiplications each) and everything above them is composed from those; the hot
// instruction footprint stays within the instruction cache while the ~22 k Montgomery multiplications of one check
// stream through it.
#pragma once
#include "fp.cuh"
#include "pairing_consts.inc"

namespace snarkv {

struct alignas(16) Fq2 { Fq c0, c1; };
struct alignas(16) Fq6 { Fq2 c0, c1, c2; };
struct alignas(16) Fq12 { Fq6 c0, c1; };

__device__ const Fq2 GAMMA1[6] = SNARKV_GAMMA1_INIT;
__device__ const Fq2 GAMMA2[6] = SNARKV_GAMMA2_INIT;
__device__ const Fq2 GAMMA3[6] = SNARKV_GAMMA3_INIT;
__device__ const Fq2 TWIST_B = SNARKV_TWIST_B_INIT;
__device__ const int8_t ATE_NAF[SNARKV_ATE_NAF_LEN] = SNARKV_ATE_NAF_INIT;

// ---- Fq2 ----------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ Fq2 fq2_zero() { return {fp_zero<FQ>(), fp_zero<FQ>()}; }
__device__ __forceinline__ Fq2 fq2_one() { return {fp_one<FQ>(), fp_zero<FQ>()}; }
__device__ __forceinline__ bool fq2_is_zero(const Fq2& a) { return fp_is_zero(a.c0) && fp_is_zero(a.c1); }
__device__ __forceinline__ bool fq2_eq(const Fq2& a, const Fq2& b) { return fp_eq(a.c0, b.c0) && fp_eq(a.c1, b.c1); }
__device__ __forceinline__ Fq2 fq2_add(const Fq2& a, const Fq2& b) { return {fp_add(a.c0, b.c0), fp_add(a.c1, b.c1)}; }
__device__ __forceinline__ Fq2 fq2_sub(const Fq2& a, const Fq2& b) { return {fp_sub(a.c0, b.c0), fp_sub(a.c1, b.c1)}; }
__device__ __forceinline__ Fq2 fq2_neg(const Fq2& a) { return {fp_neg(a.c0), fp_neg(a.c1)}; }
__device__ __forceinline__ Fq2 fq2_dbl(const Fq2& a) { return {fp_dbl(a.c0), fp_dbl(a.c1)}; }
__device__ __forceinline__ Fq2 fq2_conj(const Fq2& a) { return {a.c0, fp_neg(a.c1)}; }
// (a0 b0 - a1 b1) + (a0 b1 + a1 b0) i.  Default: two dual-product blocks (each two products under one reduction: 4 x 64 + 2 x 64
// partial products — the same as Karatsuba's 3 x 128 — but one negation instead of five modular additions / subtractions around
// them); SNARKV_FQ2_KARATSUBA restores the three-multiplication form.
static __device__ __noinline__ Fq2 fq2_mul(const Fq2& a, const Fq2& b) {
#if defined(SNARKV_FQ2_KARATSUBA) || !defined(SNARKV_PTX_FQ_MUL2ADD)
    Fq t0 = fp_mul(a.c0, b.c0);
    Fq t1 = fp_mul(a.c1, b.c1);
    Fq t2 = fp_mul(fp_add(a.c0, a.c1), fp_add(b.c0, b.c1));
    return {fp_sub(t0, t1), fp_sub(fp_sub(t2, t0), t1)};
#else
    return {fp_mul2add(a.c0, b.c0, fp_neg(a.c1), b.c1), fp_mul2add(a.c0, b.c1, a.c1, b.c0)};
#endif
}
// (c0 + c1)(c0 - c1), 2 c0 c1: 2 Montgomery multiplications
static __device__ __noinline__ Fq2 fq2_sqr(const Fq2& a) {
    Fq t = fp_mul(a.c0, a.c1);
    return {fp_mul(fp_add(a.c0, a.c1), fp_sub(a.c0, a.c1)), fp_dbl(t)};
}
__device__ __forceinline__ Fq2 fq2_scale(const Fq2& a, const Fq& k) { return {fp_mul(a.c0, k), fp_mul(a.c1, k)}; }
// * xi = (9 + i):  (9 c0 - c1) + (9 c1 + c0) i
__device__ __forceinline__ Fq2 fq2_mul_xi(const Fq2& a) {
    Fq t0 = fp_dbl(fp_dbl(fp_dbl(a.c0)));
    Fq t1 = fp_dbl(fp_dbl(fp_dbl(a.c1)));
    return {fp_sub(fp_add(t0, a.c0), a.c1), fp_add(fp_add(t1, a.c1), a.c0)};
}
static __device__ __noinline__ Fq2 fq2_inv(const Fq2& a) {
#ifdef SNARKV_TOWER_SERIAL_INV
    Fq n = fp_inv_serial(fp_add(fp_sqr(a.c0), fp_sqr(a.c1)));
#else
    Fq n = fp_inv(fp_add(fp_sqr(a.c0), fp_sqr(a.c1)));
#endif
    return {fp_mul(a.c0, n), fp_neg(fp_mul(a.c1, n))};
}

// ---- Fq6 ----------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ Fq6 fq6_zero() { return {fq2_zero(), fq2_zero(), fq2_zero()}; }
__device__ __forceinline__ Fq6 fq6_one() { return {fq2_one(), fq2_zero(), fq2_zero()}; }
__device__ __forceinline__ Fq6 fq6_add(const Fq6& a, const Fq6& b) { return {fq2_add(a.c0, b.c0), fq2_add(a.c1, b.c1), fq2_add(a.c2, b.c2)}; }
__device__ __forceinline__ Fq6 fq6_sub(const Fq6& a, const Fq6& b) { return {fq2_sub(a.c0, b.c0), fq2_sub(a.c1, b.c1), fq2_sub(a.c2, b.c2)}; }
__device__ __forceinline__ Fq6 fq6_neg(const Fq6& a) { return {fq2_neg(a.c0), fq2_neg(a.c1), fq2_neg(a.c2)}; }
__device__ __forceinline__ Fq6 fq6_mul_v(const Fq6& a) { return {fq2_mul_xi(a.c2), a.c0, a.c1}; }
// Karatsuba over Fq2: 6 Fq2 multiplications
static __device__ __noinline__ Fq6 fq6_mul(const Fq6& a, const Fq6& b) {
    Fq2 aa = fq2_mul(a.c0, b.c0), bb = fq2_mul(a.c1, b.c1), cc = fq2_mul(a.c2, b.c2);
    Fq2 t1 = fq2_mul(fq2_add(a.c1, a.c2), fq2_add(b.c1, b.c2));
    t1 = fq2_add(fq2_mul_xi(fq2_sub(fq2_sub(t1, bb), cc)), aa);
    Fq2 t2 = fq2_mul(fq2_add(a.c0, a.c1), fq2_add(b.c0, b.c1));
    t2 = fq2_add(fq2_sub(fq2_sub(t2, aa), bb), fq2_mul_xi(cc));
    Fq2 t3 = fq2_mul(fq2_add(a.c0, a.c2), fq2_add(b.c0, b.c2));
    t3 = fq2_add(fq2_sub(fq2_sub(t3, aa), cc), bb);
    return {t1, t2, t3};
}
// (c0 + c1 v + c2 v^2)(b0 + b1 v): 5 Fq2 multiplications
static __device__ __noinline__ Fq6 fq6_mul_by_01(const Fq6& a, const Fq2& b0, const Fq2& b1) {
    Fq2 a0b0 = fq2_mul(a.c0, b0), a1b1 = fq2_mul(a.c1, b1);
    Fq2 r0 = fq2_add(fq2_mul_xi(fq2_mul(a.c2, b1)), a0b0);
    Fq2 r1 = fq2_sub(fq2_sub(fq2_mul(fq2_add(a.c0, a.c1), fq2_add(b0, b1)), a0b0), a1b1);
    Fq2 r2 = fq2_add(fq2_mul(a.c2, b0), a1b1);
    return {r0, r1, r2};
}
__device__ __forceinline__ Fq6 fq6_scale(const Fq6& a, const Fq2& k) { return {fq2_mul(a.c0, k), fq2_mul(a.c1, k), fq2_mul(a.c2, k)}; }
static __device__ __noinline__ Fq6 fq6_inv(const Fq6& a) {
    Fq2 t0 = fq2_sub(fq2_sqr(a.c0), fq2_mul_xi(fq2_mul(a.c1, a.c2)));
    Fq2 t1 = fq2_sub(fq2_mul_xi(fq2_sqr(a.c2)), fq2_mul(a.c0, a.c1));
    Fq2 t2 = fq2_sub(fq2_sqr(a.c1), fq2_mul(a.c0, a.c2));
    Fq2 d = fq2_add(fq2_mul(a.c0, t0), fq2_add(fq2_mul_xi(fq2_mul(a.c2, t1)), fq2_mul_xi(fq2_mul(a.c1, t2))));
    d = fq2_inv(d);
    return {fq2_mul(t0, d), fq2_mul(t1, d), fq2_mul(t2, d)};
}

// ---- Fq12 ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ Fq12 fq12_one() { return {fq6_one(), fq6_zero()}; }
__device__ __forceinline__ Fq12 fq12_conj(const Fq12& a) { return {a.c0, fq6_neg(a.c1)}; }
static __device__ __noinline__ Fq12 fq12_mul(const Fq12& a, const Fq12& b) {
    Fq6 aa = fq6_mul(a.c0, b.c0), bb = fq6_mul(a.c1, b.c1);
    Fq6 m = fq6_mul(fq6_add(a.c0, a.c1), fq6_add(b.c0, b.c1));
    return {fq6_add(aa, fq6_mul_v(bb)), fq6_sub(fq6_sub(m, aa), bb)};
}
// complex squaring: (c0 + c1 w)^2 via (c0 + c1)(c0 + v c1) = c0^2 + v c1^2 + (1 + v) c0 c1
static __device__ __noinline__ Fq12 fq12_sqr(const Fq12& a) {
    Fq6 ab = fq6_mul(a.c0, a.c1);
    Fq6 t = fq6_mul(fq6_add(a.c0, a.c1), fq6_add(a.c0, fq6_mul_v(a.c1)));
    t = fq6_sub(fq6_sub(t, ab), fq6_mul_v(ab));
    return {t, fq6_add(ab, ab)};
}
static __device__ __noinline__ Fq12 fq12_inv(const Fq12& a) {
    Fq6 d = fq6_inv(fq6_sub(fq6_mul(a.c0, a.c0), fq6_mul_v(fq6_mul(a.c1, a.c1))));
    return {fq6_mul(a.c0, d), fq6_neg(fq6_mul(a.c1, d))};
}
// sparse product with a line  l0 + l3 w + l4 w^3  (tower slots c0.c0, c1.c0, c1.c1): 13 Fq2 multiplications
static __device__ __noinline__ Fq12 fq12_mul_by_034(const Fq12& f, const Fq2& l0, const Fq2& l3, const Fq2& l4) {
    Fq6 aA = fq6_scale(f.c0, l0);
    Fq6 bB = fq6_mul_by_01(f.c1, l3, l4);
    Fq6 m = fq6_mul_by_01(fq6_add(f.c0, f.c1), fq2_add(l0, l3), l4);
    return {fq6_add(aA, fq6_mul_v(bB)), fq6_sub(fq6_sub(m, aA), bB)};
}
// f^(p^k), k = 1..3.  In the basis 1, w, ..., w^5 over Fq2 the slot with w-power i maps to conj^k(a_i) * GAMMA_k[i];
// tower slots c0.c0, c0.c1, c0.c2, c1.c0, c1.c1, c1.c2 carry w-powers 0, 2, 4, 1, 3, 5.
static __device__ __noinline__ Fq12 fq12_frobenius(const Fq12& a, int k) {
    const Fq2* g = (k == 1) ? GAMMA1 : (k == 2) ? GAMMA2 : GAMMA3;
    const bool cj = (k & 1) != 0;
    Fq12 r;
    r.c0.c0 = cj ? fq2_conj(a.c0.c0) : a.c0.c0;
    r.c0.c1 = fq2_mul(cj ? fq2_conj(a.c0.c1) : a.c0.c1, g[2]);
    r.c0.c2 = fq2_mul(cj ? fq2_conj(a.c0.c2) : a.c0.c2, g[4]);
    r.c1.c0 = fq2_mul(cj ? fq2_conj(a.c1.c0) : a.c1.c0, g[1]);
    r.c1.c1 = fq2_mul(cj ? fq2_conj(a.c1.c1) : a.c1.c1, g[3]);
    r.c1.c2 = fq2_mul(cj ? fq2_conj(a.c1.c2) : a.c1.c2, g[5]);
    return r;
}
// Granger-Scott squaring, valid in the cyclotomic subgroup (after the easy part of the final exponentiation).
// With Fq4 = Fq2[t]/(t^2 - xi), t = w^3 and f = g0 + g1 w + g2 w^2 (g0 = (c0.c0, c1.c1), g1 = (c1.c0, c0.c2),
// g2 = (c0.c1, c1.c2)):  f^2 = (3 g0^2 - 2 conj g0) + (3 t g2^2 + 2 conj g1) w + (3 g1^2 - 2 conj g2) w^2.
__device__ __forceinline__ void fq4_sqr(const Fq2& a, const Fq2& b, Fq2& r0, Fq2& r1) {
    Fq2 a2 = fq2_sqr(a), b2 = fq2_sqr(b);
    r0 = fq2_add(fq2_mul_xi(b2), a2);
    r1 = fq2_sub(fq2_sub(fq2_sqr(fq2_add(a, b)), a2), b2);
}
static __device__ __noinline__ Fq12 fq12_cyclotomic_sqr(const Fq12& f) {
    Fq2 a0, a1, b0, b1, d0, d1;
    fq4_sqr(f.c0.c0, f.c1.c1, a0, a1);
    fq4_sqr(f.c1.c0, f.c0.c2, b0, b1);
    fq4_sqr(f.c0.c1, f.c1.c2, d0, d1);
    Fq12 r;
    r.c0.c0 = fq2_add(fq2_dbl(fq2_sub(a0, f.c0.c0)), a0);
    r.c1.c1 = fq2_add(fq2_dbl(fq2_add(a1, f.c1.c1)), a1);
    r.c0.c1 = fq2_add(fq2_dbl(fq2_sub(b0, f.c0.c1)), b0);
    r.c1.c2 = fq2_add(fq2_dbl(fq2_add(b1, f.c1.c2)), b1);
    Fq2 td1 = fq2_mul_xi(d1);
    r.c1.c0 = fq2_add(fq2_dbl(fq2_add(td1, f.c1.c0)), td1);
    r.c0.c2 = fq2_add(fq2_dbl(fq2_sub(d0, f.c0.c2)), d0);
    return r;
}
__device__ __forceinline__ bool fq12_is_one(const Fq12& f) {
    return fp_eq(f.c0.c0.c0, fp_one<FQ>()) && fp_is_zero(f.c0.c0.c1) && fq2_is_zero(f.c0.c1) && fq2_is_zero(f.c0.c2) &&
           fq2_is_zero(f.c1.c0) && fq2_is_zero(f.c1.c1) && fq2_is_zero(f.c1.c2);
}

__device__ __forceinline__ Fq2 fq2_load(const void* p) {
    const uint8_t* q = reinterpret_cast<const uint8_t*>(p);
    return {fp_load<FQ>(q), fp_load<FQ>(q + 32)};
}
__device__ __forceinline__ void fq2_store(void* p, const Fq2& a) {
    uint8_t* q = reinterpret_cast<uint8_t*>(p);
    fp_store<FQ>(q, a.c0);
    fp_store<FQ>(q + 32, a.c1);
}

}  // namespace snarkv
