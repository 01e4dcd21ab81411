// pairing_coop.cu — block-cooperative KZG decision: ONE thread block per accumulator.
//
// Same contract and same results as the one-thread-per-check kernels of pairing.cu (reference: `KzgAs::decide`,
// snark-verifier/src/pcs/kzg/decider.rs:70-82 → halo2curves `multi_miller_loop` + `final_exponentiation` + `is_identity`),
// but organised for LATENCY: a single check is what the fused batch-verification path ends in (one RLC-accumulated
// accumulator → one pairing, decider.rs:146-185), and one GPU thread needs ~13 ms for it.
//
// Fq12 is held in shared memory in the basis 1, w, ..., w^5 over Fq2 (w^6 = xi = 9 + u), 12 Fq words `[2 i + part]`.
// Every multiplicative Fq12 operation is three barrier-separated phases:
//   1. products   thread t < 12 nb computes ONE Montgomery product A[x] * B[y]            (144 threads for a full product)
//   2. columns    thread t < 22 sums the products of equal w-degree d = 0..10 into C_d (re / im part)
//   3. fold       thread t < 12 computes out_k = C_k + xi * C_{k+6}
// so the dependent chain per Fq12 operation is one multiplication + ~20 additions instead of 54 multiplications.
// Line operands of the NEXT sparse multiplication are prepared by the six spare threads 144..149 while phase 1 runs.
#include "ctx.hpp"
#include "g1.cuh"
#define SNARKV_TOWER_SERIAL_INV 1   // the one Fq inversion of a check sits on the serial path: binary-GCD inverse
#include "tower.cuh"

namespace snarkv {

namespace coop {

// threads per check: 160 (5 warps: one product per thread, lowest latency) or 32 (one warp: 4-5 products per lane, ~5x the
// checks per SM).  Template parameter NT of the kernel and of every cooperative operation below.
constexpr int NREG = 14;         // Fq12 registers in shared memory
constexpr int NUM_COEFFS = SNARKV_ATE_NUM_COEFFS;
constexpr int NAF_LEN = SNARKV_ATE_NAF_LEN;

struct Smem {
    Fq reg[NREG][12];
    Fq prod[144];
    Fq cd[22];
    Fq line[2][6];   // staged sparse operands (cy*yP).re/.im, (cx*xP).re/.im, c0.re/.im — double buffered
    Fq px[2], py[2];
    int live[2];
    int bad;
};

enum Reg { F = 0, T0, T1, FU, FU2, FU3, Y0, Y1, Y2, Y3, Y4, Y5, Y6, BASE };

__device__ __forceinline__ Fq ld(const Fq* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 lo = q[0], hi = q[1];
    Fq r;
    r.v[0] = lo.x; r.v[1] = lo.y; r.v[2] = lo.z; r.v[3] = lo.w; r.v[4] = hi.x; r.v[5] = hi.y; r.v[6] = hi.z; r.v[7] = hi.w;
    return r;
}
__device__ __forceinline__ void st(Fq* p, const Fq& a) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(a.v[0], a.v[1], a.v[2], a.v[3]);
    q[1] = make_uint4(a.v[4], a.v[5], a.v[6], a.v[7]);
}
__device__ __forceinline__ Fq times9(const Fq& a) { return fp_add(fp_dbl(fp_dbl(fp_dbl(a))), a); }

// line preparation for coefficient `idx` of pair `pair` into sm.line[buf]; executed by threads 144..149
template <int NT>
static __device__ __noinline__ void prep_line(Smem& sm, const uint8_t* __restrict__ coeffs, int pair, int idx, int buf, int t) {
    const int k = t - (NT - 6);   // the last six threads of the block
    if (k < 0) return;
    const uint8_t* c = coeffs + ((size_t)pair * NUM_COEFFS + idx) * 192;   // cy(64) | cx(64) | c0(64)
    Fq v = fp_load<FQ>(c + 32 * k);
    if (k < 2) v = fp_mul(v, ld(&sm.py[pair]));
    else if (k < 4) v = fp_mul(v, ld(&sm.px[pair]));
    st(&sm.line[buf][k], v);
}

// phases 2 + 3 shared by the full and the sparse product.  nb = 12 (full) or 6 (sparse, B support w^0, w^1, w^3).
static __device__ __noinline__ void columns_and_fold(Smem& sm, Fq* dst, int nb, int t) {
    if (t < 22) {
        const int d = t >> 1, part = t & 1;
        Fq acc = fp_zero<FQ>();
#pragma unroll 1
        for (int i = 0; i < 6; ++i) {
            const int j = d - i;
            if (j < 0 || j > 5) continue;
            int jc;
            if (nb == 12) jc = j;
            else { jc = (j == 0) ? 0 : (j == 1) ? 1 : (j == 3) ? 2 : -1; if (jc < 0) continue; }
            const Fq* pr = &sm.prod[(2 * i) * nb + 2 * jc];        // A_i.re * B_j.re , A_i.re * B_j.im
            const Fq* pi = &sm.prod[(2 * i + 1) * nb + 2 * jc];    // A_i.im * B_j.re , A_i.im * B_j.im
            if (part == 0) acc = fp_sub(fp_add(acc, ld(pr)), ld(pi + 1));
            else acc = fp_add(fp_add(acc, ld(pr + 1)), ld(pi));
        }
        st(&sm.cd[t], acc);
    }
    __syncthreads();
    if (t < 12) {
        const int k = t >> 1, part = t & 1;
        Fq r = ld(&sm.cd[t]);
        if (k + 6 <= 10) {
            const Fq hr = ld(&sm.cd[2 * (k + 6)]), hi = ld(&sm.cd[2 * (k + 6) + 1]);
            r = part == 0 ? fp_sub(fp_add(r, times9(hr)), hi) : fp_add(fp_add(r, times9(hi)), hr);
        }
        st(&dst[t], r);
    }
    __syncthreads();
}

// dst = a * b (full Fq12 product; dst may alias a or b).  While the products run, threads 144..149 optionally prepare the
// line operands (pair, idx) for a later sparse product into sm.line[buf].
template <int NT>
static __device__ __noinline__ void op_mul(Smem& sm, Fq* dst, const Fq* a, const Fq* b, int t) {
#pragma unroll 1
    for (int q = t; q < 144; q += NT) st(&sm.prod[q], fp_mul(ld(&a[q / 12]), ld(&b[q % 12])));
    __syncthreads();
    columns_and_fold(sm, dst, 12, t);
}
// dst = a * line[buf]  (sparse: B = l0 + l1 w + l3 w^3)
template <int NT>
static __device__ __noinline__ void op_sparse(Smem& sm, Fq* dst, const Fq* a, int buf, int t) {
#pragma unroll 1
    for (int q = t; q < 72; q += NT) st(&sm.prod[q], fp_mul(ld(&a[q / 6]), ld(&sm.line[buf][q % 6])));
    __syncthreads();
    columns_and_fold(sm, dst, 6, t);
}
// dst = conj^k(a) coefficient-wise times GAMMA_k  (Frobenius f^(p^k), k = 1..3; w-basis index = w-power)
static __device__ __noinline__ void op_frobenius(Smem& sm, Fq* dst, const Fq* a, int k, int t) {
    const Fq2* g = (k == 1) ? GAMMA1 : (k == 2) ? GAMMA2 : GAMMA3;
    if (t < 24) {
        const int i = t >> 2, s = (t >> 1) & 1, tt = t & 1;
        const Fq gv = tt ? g[i].c1 : g[i].c0;
        st(&sm.prod[t], fp_mul(ld(&a[2 * i + s]), gv));
    }
    __syncthreads();
    if (t < 12) {
        const int i = t >> 1, part = t & 1;
        const Fq* p = &sm.prod[4 * i];   // [re*g0, re*g1, im*g0, im*g1]
        Fq r;
        if (k & 1) r = part == 0 ? fp_add(ld(p), ld(p + 3)) : fp_sub(ld(p + 1), ld(p + 2));   // conjugated input: im -> -im
        else r = part == 0 ? fp_sub(ld(p), ld(p + 3)) : fp_add(ld(p + 1), ld(p + 2));
        st(&dst[t], r);
    }
    __syncthreads();
}
// f^(p^6): negate the odd powers of w
__device__ __forceinline__ void op_conj(Smem&, Fq* dst, const Fq* a, int t) {
    if (t < 12) {
        const int i = t >> 1;
        Fq v = ld(&a[t]);
        st(&dst[t], (i & 1) ? fp_neg(v) : v);
    }
    __syncthreads();
}
__device__ __forceinline__ void op_copy(Fq* dst, const Fq* a, int t) {
    if (t < 12) st(&dst[t], ld(&a[t]));
    __syncthreads();
}

// tower slot m (c0.c0.c0, c0.c0.c1, c0.c1.c0, ... serialisation order) -> w-basis word
__device__ __forceinline__ int tower_to_w(int m) {
    const int slot = m >> 1, part = m & 1;
    const int wp = (slot < 3) ? 2 * slot : 2 * (slot - 3) + 1;
    return 2 * wp + part;
}

// dst = a^-1 by thread 0 with the tower formulas (one Fq inversion inside); the only serial piece of a check
static __device__ __noinline__ void op_inverse_serial(Fq* dst, const Fq* a, int t) {
    if (t == 0) {
        Fq12 x;
        Fq* xs = reinterpret_cast<Fq*>(&x);
        for (int m = 0; m < 12; ++m) xs[m] = ld(&a[tower_to_w(m)]);
        Fq12 xi = fq12_inv(x);
        const Fq* is = reinterpret_cast<const Fq*>(&xi);
        for (int m = 0; m < 12; ++m) st(&dst[tower_to_w(m)], is[m]);
    }
    __syncthreads();
}

// dst = a^u (u = BN parameter, 63 bits), a in the cyclotomic subgroup.  Uses BASE and dst as scratch.
template <int NT>
static __device__ __noinline__ void op_exp_by_u(Smem& sm, Fq* dst, const Fq* a, int t) {
    Fq* base = sm.reg[BASE];
    op_copy(base, a, t);
    op_copy(dst, a, t);
    const uint64_t u = SNARKV_BN_U;
#pragma unroll 1
    for (int i = 61; i >= 0; --i) {
        op_mul<NT>(sm, dst, dst, dst, t);
        if ((u >> i) & 1ull) op_mul<NT>(sm, dst, dst, base, t);
    }
}

}  // namespace coop

using namespace coop;

template <int NT>
__global__ void __launch_bounds__(NT, (NT == 32) ? 16 : 3) k_kzg_decide_coop(const uint8_t* __restrict__ lhs, const uint8_t* __restrict__ rhs, size_t N, int format,
                                                             const uint8_t* __restrict__ coeffs, const int* __restrict__ infinity,
                                                             uint8_t* __restrict__ accept, uint8_t* __restrict__ gt_out) {
    __shared__ Smem sm;
    const int t = threadIdx.x;
    for (size_t chk = blockIdx.x; chk < N; chk += gridDim.x) {
        // ---- load + validate the two G1 points (threads 0 and 1), f = 1 ------------------------------------------------
        if (t == 0) sm.bad = 0;
        __syncthreads();
        if (t < 2) {
            const uint8_t* src = (t == 0) ? lhs : rhs;
            G1Affine p = g1_affine_load(src, chk);
            bool ok = true;
            if (format == SNARKV_CANONICAL) {
                if (!fp_is_canonical(p.x) || !fp_is_canonical(p.y)) ok = false;
                p.x = fp_to_mont(p.x);
                p.y = fp_to_mont(p.y);
            }
            if (!g1_affine_is_on_curve(p)) ok = false;
            st(&sm.px[t], p.x);
            st(&sm.py[t], p.y);
            sm.live[t] = (ok && !g1_affine_is_identity(p) && !infinity[t]) ? 1 : 0;
            if (!ok) sm.bad = 1;
        }
        if (t < 12) st(&sm.reg[F][t], t == 0 ? fp_one<FQ>() : fp_zero<FQ>());
        __syncthreads();
        if (sm.bad) {   // uniform: rejected input (`from_xy` would have failed, accumulator.rs:75-78)
            if (t == 0) accept[chk] = 0;
            if (gt_out && t < 12) fp_store<FQ>(gt_out + chk * 384 + 32 * t, fp_zero<FQ>());
            __syncthreads();
            continue;
        }
        const int live0 = sm.live[0], live1 = sm.live[1];
        Fq* f = sm.reg[F];

        // ---- multi-Miller loop (shared squarings; pairs with an identity skipped) ------------------------------------------
        // line operands are staged one step ahead by threads 144..149 during the previous operation's product phase
        int idx = 0;
        if (live0) prep_line<NT>(sm, coeffs, 0, 0, 0, t);
        if (live1) prep_line<NT>(sm, coeffs, 1, 0, 1, t);
        __syncthreads();
#pragma unroll 1
        for (int b = NAF_LEN - 2; b >= 0; --b) {
            const int steps = (ATE_NAF[b] != 0) ? 2 : 1;
            if (b != NAF_LEN - 2) op_mul<NT>(sm, f, f, f, t);
#pragma unroll 1
            for (int s = 0; s < steps; ++s) {
                // buffers 0/1 hold the operands for (pair 0, idx) / (pair 1, idx); refill each buffer for idx + 1 right after use
                if (live0) op_sparse<NT>(sm, f, f, 0, t);
                if (live1) op_sparse<NT>(sm, f, f, 1, t);
                ++idx;
                if (idx < NUM_COEFFS) {
                    if (live0) prep_line<NT>(sm, coeffs, 0, idx, 0, t);
                    if (live1) prep_line<NT>(sm, coeffs, 1, idx, 1, t);
                    __syncthreads();
                }
            }
        }
#pragma unroll 1
        for (int extra = 0; extra < 2; ++extra) {
            if (live0) op_sparse<NT>(sm, f, f, 0, t);
            if (live1) op_sparse<NT>(sm, f, f, 1, t);
            ++idx;
            if (idx < NUM_COEFFS) {
                if (live0) prep_line<NT>(sm, coeffs, 0, idx, 0, t);
                if (live1) prep_line<NT>(sm, coeffs, 1, idx, 1, t);
                __syncthreads();
            }
        }

        // ---- final exponentiation -----------------------------------------------------------------------------------------
        // easy part: f <- conj(f) * f^-1 ; f <- f^(p^2) * f.   The single Fq12 inversion is serial (thread 0, tower code).
        op_inverse_serial(sm.reg[T0], f, t);
        op_conj(sm, sm.reg[T1], f, t);
        op_mul<NT>(sm, f, sm.reg[T1], sm.reg[T0], t);
        op_frobenius(sm, sm.reg[T0], f, 2, t);
        op_mul<NT>(sm, f, sm.reg[T0], f, t);
        // hard part (p^4 - p^2 + 1)/r: y0 y1^2 y2^6 y3^12 y4^18 y5^30 y6^36 (Devegili-Scott-Dahab)
        op_exp_by_u<NT>(sm, sm.reg[FU], f, t);
        op_exp_by_u<NT>(sm, sm.reg[FU2], sm.reg[FU], t);
        op_exp_by_u<NT>(sm, sm.reg[FU3], sm.reg[FU2], t);
        op_frobenius(sm, sm.reg[Y0], f, 1, t);                       // y0 = f^p f^(p^2) f^(p^3)
        op_frobenius(sm, sm.reg[T0], f, 2, t);
        op_mul<NT>(sm, sm.reg[Y0], sm.reg[Y0], sm.reg[T0], t);
        op_frobenius(sm, sm.reg[T0], f, 3, t);
        op_mul<NT>(sm, sm.reg[Y0], sm.reg[Y0], sm.reg[T0], t);
        op_conj(sm, sm.reg[Y1], f, t);                               // y1 = 1/f
        op_frobenius(sm, sm.reg[Y2], sm.reg[FU2], 2, t);             // y2 = (f^(u^2))^(p^2)
        op_frobenius(sm, sm.reg[T0], sm.reg[FU], 1, t);              // y3 = 1/(f^u)^p
        op_conj(sm, sm.reg[Y3], sm.reg[T0], t);
        op_frobenius(sm, sm.reg[T0], sm.reg[FU2], 1, t);             // y4 = 1/(f^u (f^(u^2))^p)
        op_mul<NT>(sm, sm.reg[T0], sm.reg[T0], sm.reg[FU], t);
        op_conj(sm, sm.reg[Y4], sm.reg[T0], t);
        op_conj(sm, sm.reg[Y5], sm.reg[FU2], t);                     // y5 = 1/f^(u^2)
        op_frobenius(sm, sm.reg[T0], sm.reg[FU3], 1, t);             // y6 = 1/(f^(u^3) (f^(u^3))^p)
        op_mul<NT>(sm, sm.reg[T0], sm.reg[T0], sm.reg[FU3], t);
        op_conj(sm, sm.reg[Y6], sm.reg[T0], t);
        Fq* t0 = sm.reg[T0];
        Fq* t1 = sm.reg[T1];
        op_mul<NT>(sm, t0, sm.reg[Y6], sm.reg[Y6], t);                   // t0 = y6^2 y4 y5
        op_mul<NT>(sm, t0, t0, sm.reg[Y4], t);
        op_mul<NT>(sm, t0, t0, sm.reg[Y5], t);
        op_mul<NT>(sm, t1, sm.reg[Y3], sm.reg[Y5], t);                   // t1 = y3 y5 t0
        op_mul<NT>(sm, t1, t1, t0, t);
        op_mul<NT>(sm, t0, t0, sm.reg[Y2], t);                           // t0 = t0 y2
        op_mul<NT>(sm, t1, t1, t1, t);                                   // t1 = (t1^2 t0)^2
        op_mul<NT>(sm, t1, t1, t0, t);
        op_mul<NT>(sm, t1, t1, t1, t);
        op_mul<NT>(sm, t0, t1, sm.reg[Y1], t);                           // t0 = t1 y1
        op_mul<NT>(sm, t1, t1, sm.reg[Y0], t);                           // t1 = t1 y0
        op_mul<NT>(sm, t0, t0, t0, t);                                   // gt = t0^2 t1
        op_mul<NT>(sm, f, t0, t1, t);

        // ---- verdict + optional GT bytes ------------------------------------------------------------------------------------
        if (t == 0) {
            bool one = fp_eq(ld(&f[0]), fp_one<FQ>());
            for (int m = 1; m < 12; ++m) one = one && fp_is_zero(ld(&f[m]));
            accept[chk] = one ? 1 : 0;
        }
        if (gt_out && t < 12) fp_store<FQ>(gt_out + chk * 384 + 32 * t, fp_from_mont(ld(&f[tower_to_w(t)])));
        __syncthreads();
    }
}

int kzg_decide_coop_device(snarkv_ctx* ctx, const void* d_lhs, const void* d_rhs, size_t N, int format, void* d_accept, void* d_gt) {
    const uint8_t* base = (const uint8_t*)ctx->d_key_coeffs;
    const int* d_inf = (const int*)(base + (size_t)2 * coop::NUM_COEFFS * 192);
    // up to one wave of 160-thread blocks (3 per SM): lowest latency; beyond that one warp per check (16 per SM)
    const bool wide = ctx->pairing_mode == 3 || (ctx->pairing_mode != 4 && N <= (size_t)ctx->sm_count * 3);
    const size_t cap = (size_t)ctx->sm_count * (wide ? 3 : 16) * 4;
    const unsigned blocks = (unsigned)(N < cap ? N : cap);
    Stage sg(ctx, wide ? "kzg_decide_block_per_check" : "kzg_decide_warp_per_check");
    if (wide)
        k_kzg_decide_coop<160><<<blocks, 160, 0, ctx->stream>>>((const uint8_t*)d_lhs, (const uint8_t*)d_rhs, N, format, base, d_inf,
                                                              (uint8_t*)d_accept, (uint8_t*)d_gt);
    else
        k_kzg_decide_coop<32><<<blocks, 32, 0, ctx->stream>>>((const uint8_t*)d_lhs, (const uint8_t*)d_rhs, N, format, base, d_inf,
                                                            (uint8_t*)d_accept, (uint8_t*)d_gt);
    SNARKV_LAUNCH_CHECK(ctx, "k_kzg_decide_coop");
    sg.launched();
    return SNARKV_OK;
}

}  // namespace snarkv
