// g1.cuh — BN254 G1 (y^2 = x^3 + 3 over Fq, a = 0) point arithmetic for the MSM kernels.
//
// Device-side replacement for halo2curves 0.6.0 `bn256::{G1, G1Affine}` as used at the reference's call sites
// snark-verifier/src/loader/native.rs:67-70 (`*base * scalar`, `acc + value`, `.to_affine()`) and
// util/msm.rs:236-256,286,298-302 (bucket `+=`, `double`, running sums).
//
// Accumulators use extended Jacobian "XYZZ" coordinates (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2): the mixed addition
// costs 8M + 2S = 10 Montgomery multiplications versus 11 for Jacobian madd-2007-bl, and needs no field halving.
// Identity: ZZ == 0.  Affine identity: (0, 0) — halo2curves' encoding.  Results are canonical after to_affine, so the
// coordinate system is invisible at the C-ABI boundary.
#pragma once
#include "fp.cuh"

namespace snarkv {

struct alignas(16) G1Affine {
    Fq x, y;
};
struct alignas(16) G1Xyzz {
    Fq x, y, zz, zzz;
};
struct alignas(16) G1Jac {  // halo2curves' G1 layout: (X, Y, Z) with x = X/Z^2, y = Y/Z^3; identity Z == 0
    Fq x, y, z;
};

__device__ __forceinline__ bool g1_affine_is_identity(const G1Affine& p) { return fp_is_zero(p.x) && fp_is_zero(p.y); }
__device__ __forceinline__ bool xyzz_is_identity(const G1Xyzz& p) { return fp_is_zero(p.zz); }

__device__ __forceinline__ G1Xyzz xyzz_identity() {
    G1Xyzz r;
    r.x = fp_zero<FQ>(); r.y = fp_one<FQ>(); r.zz = fp_zero<FQ>(); r.zzz = fp_zero<FQ>();
    return r;
}
__device__ __forceinline__ G1Xyzz xyzz_from_affine(const G1Affine& p) {
    if (g1_affine_is_identity(p)) return xyzz_identity();
    G1Xyzz r;
    r.x = p.x; r.y = p.y; r.zz = fp_one<FQ>(); r.zzz = fp_one<FQ>();
    return r;
}

__device__ __forceinline__ G1Affine g1_affine_load(const void* base, size_t idx) {
    const uint8_t* p = reinterpret_cast<const uint8_t*>(base) + idx * 64;
    G1Affine r;
    r.x = fp_load<FQ>(p);
    r.y = fp_load<FQ>(p + 32);
    return r;
}
__device__ __forceinline__ void g1_affine_store(void* base, size_t idx, const G1Affine& a) {
    uint8_t* p = reinterpret_cast<uint8_t*>(base) + idx * 64;
    fp_store<FQ>(p, a.x);
    fp_store<FQ>(p + 32, a.y);
}
__device__ __forceinline__ G1Xyzz xyzz_load(const void* base, size_t idx) {
    const uint8_t* p = reinterpret_cast<const uint8_t*>(base) + idx * 128;
    G1Xyzz r;
    r.x = fp_load<FQ>(p); r.y = fp_load<FQ>(p + 32); r.zz = fp_load<FQ>(p + 64); r.zzz = fp_load<FQ>(p + 96);
    return r;
}
__device__ __forceinline__ void xyzz_store(void* base, size_t idx, const G1Xyzz& a) {
    uint8_t* p = reinterpret_cast<uint8_t*>(base) + idx * 128;
    fp_store<FQ>(p, a.x); fp_store<FQ>(p + 32, a.y); fp_store<FQ>(p + 64, a.zz); fp_store<FQ>(p + 96, a.zzz);
}

// 2 * (x, y) for an affine, non-identity point (mdbl-2008-s-1 with a = 0)
static __device__ __noinline__ G1Xyzz xyzz_dbl_affine(const Fq& x, const Fq& y) {
    G1Xyzz r;
    if (fp_is_zero(y)) return xyzz_identity();  // order-2 points do not exist on BN254 G1; kept for totality
    Fq u = fp_dbl(y);
    Fq v = fp_sqr(u);
    Fq w = fp_mul(u, v);
    Fq s = fp_mul(x, v);
    Fq xx = fp_sqr(x);
    Fq m = fp_add(fp_dbl(xx), xx);
    r.x = fp_sub(fp_sqr(m), fp_dbl(s));
    r.y = fp_mul2add(m, fp_sub(s, r.x), fp_neg(w), y);   // m (s - x3) - w y under one reduction
    r.zz = v;
    r.zzz = w;
    return r;
}

// 2 * P (dbl-2008-s-1, a = 0): 6M + 3S
__device__ __forceinline__ G1Xyzz xyzz_dbl(const G1Xyzz& p) {
    if (xyzz_is_identity(p)) return p;
    G1Xyzz r;
    Fq u = fp_dbl(p.y);
    Fq v = fp_sqr(u);
    Fq w = fp_mul(u, v);
    Fq s = fp_mul(p.x, v);
    Fq xx = fp_sqr(p.x);
    Fq m = fp_add(fp_dbl(xx), xx);
    r.x = fp_sub(fp_sqr(m), fp_dbl(s));
    r.y = fp_mul2add(m, fp_sub(s, r.x), fp_neg(w), p.y);
    r.zz = fp_mul(v, p.zz);
    r.zzz = fp_mul(w, p.zzz);
    return r;
}

// acc += (x2, y2), (x2, y2) affine and NOT the identity (madd-2008-s): 8M + 2S on the common path; the two squarings are the
// dedicated SQR block and y3 is one dual-product block (fp_mul2add), i.e. 8.3 multiplications' worth of partial products.
// Exceptional cases (empty accumulator, equal points, opposite points) are handled exactly.
__device__ __forceinline__ void xyzz_madd(G1Xyzz& acc, const Fq& x2, const Fq& y2) {
    if (xyzz_is_identity(acc)) {
        acc.x = x2; acc.y = y2; acc.zz = fp_one<FQ>(); acc.zzz = fp_one<FQ>();
        return;
    }
    Fq u2 = fp_mul(x2, acc.zz);
    Fq s2 = fp_mul(y2, acc.zzz);
    Fq p = fp_sub(u2, acc.x);
    Fq r = fp_sub(s2, acc.y);
    if (fp_is_zero(p)) {
        if (fp_is_zero(r)) acc = xyzz_dbl_affine(x2, y2);
        else acc = xyzz_identity();
        return;
    }
    Fq pp = fp_sqr(p);
    Fq ppp = fp_mul(p, pp);
    Fq q = fp_mul(acc.x, pp);
    Fq x3 = fp_sub(fp_sub(fp_sqr(r), ppp), fp_dbl(q));
    acc.y = fp_mul2add(r, fp_sub(q, x3), fp_neg(acc.y), ppp);   // r (q - x3) - y1 ppp: two products, one reduction
    acc.x = x3;
    acc.zz = fp_mul(acc.zz, pp);
    acc.zzz = fp_mul(acc.zzz, ppp);
}
__device__ __forceinline__ void xyzz_madd(G1Xyzz& acc, const G1Affine& p) {
    if (g1_affine_is_identity(p)) return;
    xyzz_madd(acc, p.x, p.y);
}

// a + b, both XYZZ (add-2008-s): 12M + 2S
__device__ __forceinline__ G1Xyzz xyzz_add(const G1Xyzz& a, const G1Xyzz& b) {
    if (xyzz_is_identity(a)) return b;
    if (xyzz_is_identity(b)) return a;
    Fq u1 = fp_mul(a.x, b.zz);
    Fq u2 = fp_mul(b.x, a.zz);
    Fq s1 = fp_mul(a.y, b.zzz);
    Fq s2 = fp_mul(b.y, a.zzz);
    Fq p = fp_sub(u2, u1);
    Fq r = fp_sub(s2, s1);
    if (fp_is_zero(p)) {
        if (fp_is_zero(r)) return xyzz_dbl(a);
        return xyzz_identity();
    }
    Fq pp = fp_sqr(p);
    Fq ppp = fp_mul(p, pp);
    Fq q = fp_mul(u1, pp);
    G1Xyzz o;
    o.x = fp_sub(fp_sub(fp_sqr(r), ppp), fp_dbl(q));
    o.y = fp_mul2add(r, fp_sub(q, o.x), fp_neg(s1), ppp);
    o.zz = fp_mul(fp_mul(a.zz, b.zz), pp);
    o.zzz = fp_mul(fp_mul(a.zzz, b.zzz), ppp);
    return o;
}

// XYZZ -> Jacobian without inversion: Z = ZZZ  =>  Z^2 = ZZ^3, Z^3 = ZZZ^3  =>  (X ZZ^2, Y ZZZ^2, ZZZ)
__device__ __forceinline__ G1Jac xyzz_to_jacobian(const G1Xyzz& p) {
    G1Jac r;
    if (xyzz_is_identity(p)) {
        r.x = fp_zero<FQ>(); r.y = fp_one<FQ>(); r.z = fp_zero<FQ>();
        return r;
    }
    r.x = fp_mul(p.x, fp_sqr(p.zz));
    r.y = fp_mul(p.y, fp_sqr(p.zzz));
    r.z = p.zzz;
    return r;
}
__device__ __forceinline__ G1Xyzz jacobian_to_xyzz(const G1Jac& p) {
    if (fp_is_zero(p.z)) return xyzz_identity();
    G1Xyzz r;
    r.x = p.x; r.y = p.y;
    r.zz = fp_sqr(p.z);
    r.zzz = fp_mul(r.zz, p.z);
    return r;
}
// `Curve::to_affine` (native.rs:70): one inversion.  (0,0) for the identity.
__device__ __forceinline__ G1Affine xyzz_to_affine(const G1Xyzz& p) {
    G1Affine r;
    if (xyzz_is_identity(p)) {
        r.x = fp_zero<FQ>(); r.y = fp_zero<FQ>();
        return r;
    }
    // 1/ZZZ gives both: 1/ZZ = ZZ^2 / ZZZ^2 ... simpler: inv(ZZ * ZZZ) then split
    Fq t = fp_inv(fp_mul(p.zz, p.zzz));
    Fq izz = fp_mul(t, p.zzz);
    Fq izzz = fp_mul(t, p.zz);
    r.x = fp_mul(p.x, izz);
    r.y = fp_mul(p.y, izzz);
    return r;
}

// ---------------------------------------------------------------------------------------------------------------------
// 4-lane point arithmetic for serial critical paths (the window-combining Horner chain of util/msm.rs:285-287 and the fold
// of per-GPU partials, util/msm.rs:333-335).  Lanes 0..3 of a warp hold IDENTICAL copies of the operands; the independent
// field multiplications of each formula are computed one per lane and exchanged with shuffles, so a doubling is 3 and an
// addition 4 dependent multiplications deep instead of 9 and 14.  ONE warp calls these together; every aligned group of four lanes computes
// the same thing (`lane` = threadIdx.x & 3), lanes 0..3 publish through the 4-entry shared scratch `xch`.
// ---------------------------------------------------------------------------------------------------------------------
// exchange: lane k (< 4) publishes its product in shared memory, everyone reads all four (cheaper than shuffles here: the
// compiler wraps every partial-convergence shuffle in a WARPSYNC/ENDCOLLECTIVE pair, ~750 cycles per 8-word broadcast).
struct Fq4 { Fq v[4]; };
__device__ __forceinline__ Fq4 fq_exchange4(Fq* xch, const Fq& mine) {
    __syncwarp();
    if (threadIdx.x < 4) fp_store<FQ>(&xch[threadIdx.x], mine);
    __syncwarp();
    Fq4 o;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint4* q = reinterpret_cast<const uint4*>(&xch[k]);
        uint4 lo = q[0], hi = q[1];
        o.v[k].v[0] = lo.x; o.v[k].v[1] = lo.y; o.v[k].v[2] = lo.z; o.v[k].v[3] = lo.w;
        o.v[k].v[4] = hi.x; o.v[k].v[5] = hi.y; o.v[k].v[6] = hi.z; o.v[k].v[7] = hi.w;
    }
    return o;
}
// branch-free 4-way select (lane-dependent ternaries would compile into divergent branch regions, one per limb)
__device__ __forceinline__ Fq fq_sel4(int lane, const Fq& a0, const Fq& a1, const Fq& a2, const Fq& a3) {
    const uint32_t m0 = 0u - (uint32_t)(lane == 0), m1 = 0u - (uint32_t)(lane == 1), m2 = 0u - (uint32_t)(lane == 2),
                   m3 = 0u - (uint32_t)(lane == 3);
    Fq r;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.v[i] = (a0.v[i] & m0) | (a1.v[i] & m1) | (a2.v[i] & m2) | (a3.v[i] & m3);
    return r;
}

static __device__ __noinline__ G1Xyzz xyzz_dbl_x4(const G1Xyzz& p, int lane, Fq* xch) {
    if (xyzz_is_identity(p)) return p;
    const Fq u = fp_dbl(p.y);
    Fq4 r = fq_exchange4(xch, fp_mul(fq_sel4(lane, u, p.x, u, p.x), fq_sel4(lane, u, p.x, u, p.x)));      // V = U^2 | XX = X^2
    const Fq v = r.v[0];
    const Fq m = fp_add(fp_dbl(r.v[1]), r.v[1]);
    r = fq_exchange4(xch, fp_mul(fq_sel4(lane, u, p.x, m, v), fq_sel4(lane, v, v, m, p.zz)));             // W | S | M^2 | ZZ3
    const Fq w = r.v[0], s = r.v[1];
    G1Xyzz o;
    o.zz = r.v[3];
    o.x = fp_sub(r.v[2], fp_dbl(s));
    r = fq_exchange4(xch, fp_mul(fq_sel4(lane, m, w, w, w), fq_sel4(lane, fp_sub(s, o.x), p.y, p.zzz, p.zzz)));  // M(S-X3) | W Y | W ZZZ
    o.y = fp_sub(r.v[0], r.v[1]);
    o.zzz = r.v[2];
    return o;
}
static __device__ __noinline__ G1Xyzz xyzz_add_x4(const G1Xyzz& a, const G1Xyzz& b, int lane, Fq* xch) {
    if (xyzz_is_identity(a)) return b;
    if (xyzz_is_identity(b)) return a;
    Fq4 r = fq_exchange4(xch, fp_mul(fq_sel4(lane, a.x, b.x, a.y, b.y), fq_sel4(lane, b.zz, a.zz, b.zzz, a.zzz)));  // U1 | U2 | S1 | S2
    const Fq u1 = r.v[0], s1 = r.v[2];
    const Fq p = fp_sub(r.v[1], u1), rr = fp_sub(r.v[3], s1);
    if (fp_is_zero(p)) {
        if (fp_is_zero(rr)) return xyzz_dbl_x4(a, lane, xch);
        return xyzz_identity();
    }
    r = fq_exchange4(xch, fp_mul(fq_sel4(lane, p, rr, a.zz, a.zzz), fq_sel4(lane, p, rr, b.zz, b.zzz)));           // PP | R^2 | ZZ1 ZZ2 | ZZZ1 ZZZ2
    const Fq pp = r.v[0], r2 = r.v[1], zz12 = r.v[2], zzz12 = r.v[3];
    r = fq_exchange4(xch, fp_mul(fq_sel4(lane, p, u1, zz12, zz12), pp));                                          // PPP | Q | ZZ3
    const Fq ppp = r.v[0], q = r.v[1];
    G1Xyzz o;
    o.zz = r.v[2];
    o.x = fp_sub(fp_sub(r2, ppp), fp_dbl(q));
    r = fq_exchange4(xch, fp_mul(fq_sel4(lane, rr, s1, zzz12, zzz12), fq_sel4(lane, fp_sub(q, o.x), ppp, ppp, ppp)));  // R(Q-X3) | S1 PPP | ZZZ3
    o.y = fp_sub(r.v[0], r.v[1]);
    o.zzz = r.v[2];
    return o;
}
// to_affine with the serial (binary-GCD) inverse; every lane computes the same value
static __device__ __noinline__ G1Affine xyzz_to_affine_serial(const G1Xyzz& p) {
    G1Affine r;
    if (xyzz_is_identity(p)) {
        r.x = fp_zero<FQ>(); r.y = fp_zero<FQ>();
        return r;
    }
    Fq t = fp_inv_serial(fp_mul(p.zz, p.zzz));
    r.x = fp_mul(p.x, fp_mul(t, p.zzz));
    r.y = fp_mul(p.y, fp_mul(t, p.zz));
    return r;
}

// y^2 == x^3 + b, b = SNARKV_CURVE_B (Montgomery-form inputs); the identity (0,0) is accepted — `CurveAffine::from_xy` semantics plus
// halo2curves' identity encoding
__device__ __forceinline__ bool g1_affine_is_on_curve(const G1Affine& p) {
    if (g1_affine_is_identity(p)) return true;
    Fq b = fp_one<FQ>();                     // the curve constant: 3 (BN254 G1) or 5 (Pallas), fp.cuh
#pragma unroll
    for (int k = 1; k < SNARKV_CURVE_B; ++k) b = fp_add(b, fp_one<FQ>());
    Fq rhs = fp_add(fp_mul(fp_sqr(p.x), p.x), b);
    return fp_eq(fp_sqr(p.y), rhs);
}

// k * P for a small unsigned k (bucket-segment weights), double-and-add MSB first
static __device__ __noinline__ G1Xyzz xyzz_mul_small(const G1Xyzz& p, uint32_t k) {
    G1Xyzz acc = xyzz_identity();
    if (k == 0) return acc;
    for (int i = 31 - __clz(k); i >= 0; --i) {
        acc = xyzz_dbl(acc);
        if ((k >> i) & 1u) acc = xyzz_add(acc, p);
    }
    return acc;
}

}  // namespace snarkv
