// pairing_fast.cu — KZG decision organised for the LATENCY of one check: one 256-thread block per accumulator.
//
// Same contract and the same results as the other decision kernels (reference: `KzgAs::decide`,
// snark-verifier/src/pcs/kzg/decider.rs:70-82 → halo2curves `multi_miller_loop` + `final_exponentiation` + `is_identity`).
// Every batched verification path of this library ends in ONE pairing (RLC `decide_all`, decider.rs:146-185; the fused PLONK
// batch; `KzgAs::verify` + decide), so the latency of a single check is what those paths pay.  The first cooperative kernel
// (pairing_coop.cu) needs 2.15 ms per check: 539 barrier-separated Fq12 operations of ~4 us each, of which the Montgomery
// products are under a quarter — the rest is a 22-thread column phase and a 12-thread fold phase of ~18 dependent modular
// additions.  This kernel removes both the operations and the serial additions:
//
//   * Fq12 = Fq2[w]/(w^6 - xi), 12 Fq words `[2 i + part]` in shared memory.  A product is ONE phase: output word o = (k, part) owns a
//     group of 16 lanes; lane s < 12 multiplies a_i.{re|im} (i = s >> 1) by the matching component of b_j, j = k - i mod 6, with
//     xi folded into the OPERAND when i + j >= 6 (xi b_j = (9 re - im, 9 im + re), read from a precomputed copy when the
//     caller has one); the 12 products of a word are summed by an xor-butterfly of shuffles whose first level subtracts for real
//     words.  Critical path per product: one Montgomery multiplication + 4 shuffle/add levels, two barriers.
//   * both G2 points are fixed per deciding key, so the two sparse line factors of a Miller step are merged into one dense
//     factor whose 9 Fq2 constants (cy1 cy2, cy1 cx2, ..., xi c01 c02) are tabulated once per key (k_pair_tables); per check and
//     step the merged line costs 16 products by per-check scalars (y1 y2, y1 x2, x1 y2, x1 x2, y1, y2, x1, x2), and ALL 88 are
//     evaluated into shared memory before the loop (one warp per step).  Miller loop: 64 squarings + 88 products = 152
//     operations instead of 240.
//   * exponentiation by u in signed digits (the inverse of a cyclotomic element is its conjugate): 62 squarings + 23 products;
//   * the one Fq12 inversion by norms (a conj(a) in Fq6, then the Fq6/Fq2 norm via two Frobenius maps): 5 cooperative products
//     + one Fq inversion instead of ~100 serial multiplications.
//
// The index arithmetic of this file (term map, Frobenius signs, line slots, digit chains) is emulated thread by thread over
// exact integers by the test-side model pairing_fast_model.py and checked against the independent pairing model (tests/test_pairing_fast_model.py).
#include "ctx.hpp"
#include "g1.cuh"
#define SNARKV_TOWER_SERIAL_INV 1
#include "tower.cuh"

namespace snarkv {

namespace fast {

constexpr int NT = 256;           // threads per check
constexpr int NPROD = 192;        // 12 output words x 16 lanes
constexpr int NSTEP = SNARKV_ATE_NUM_COEFFS;
constexpr int NAF_LEN = SNARKV_ATE_NAF_LEN;
constexpr int NSLOT = 16;         // Fq2 slots of a merged-line table row (32 lanes = slot x part)

__device__ const int8_t U_NAF[SNARKV_U_NAF_LEN] = SNARKV_U_NAF_INIT;

enum Reg { F = 0, T0, T1, T2, FU, FU2, FU3, Y0, Y1, Y2, Y3, Y4, Y5, Y6, BASE, BASEX, BASEC, BASECX, NREG };

struct Smem {
    Fq reg[NREG][12];
    Fq S[NSLOT];            // per-check scalars of the merged lines
    Fq px[2], py[2];
    int live[2];
    int bad;
    int pad;
    Fq line[NSTEP][2][12];  // [step][0 = L | 1 = xi L]
};

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
__device__ __forceinline__ Fq shfl_xor(const Fq& a, int mask) {
    Fq r;
#pragma unroll
    for (int k = 0; k < 8; ++k) r.v[k] = __shfl_xor_sync(0xffffffffu, a.v[k], mask);
    return r;
}
__device__ __forceinline__ Fq times9(const Fq& a) { return fp_add(fp_dbl(fp_dbl(fp_dbl(a))), a); }

// dst = a * b.  `bx` = xi * b word by word if the caller keeps one (lines, the base of an exponentiation), else nullptr and the
// lanes that need it compute it.  dst may alias a or b.  All NT threads call; threads >= NPROD only take part in the barriers.
static __device__ __noinline__ void op_mul(Fq* dst, const Fq* a, const Fq* b, const Fq* bx, int t) {
    const int o = t >> 4, s = t & 15, k = o >> 1, part = o & 1, i = s >> 1, odd = s & 1;
    const bool act = t < NPROD && s < 12;
    const int jj = k - i;
    const bool high = jj < 0;
    const int j = high ? jj + 6 : jj;
    const int c = part ^ odd;
    const bool own_xi = high && bx == nullptr;
    Fq A = fp_zero<FQ>(), M = fp_zero<FQ>(), N = fp_zero<FQ>();
    if (act) {
        A = ld(&a[2 * i + odd]);
        M = ld(high && bx ? &bx[2 * j + c] : &b[2 * j + c]);
        if (own_xi) N = ld(&b[2 * j + (c ^ 1)]);
    }
    __syncthreads();   // every operand word is in registers: dst may be overwritten
    if (t < NPROD) {   // whole warps
        Fq v = fp_zero<FQ>();
        if (act) {
            if (own_xi) {
                const Fq m9 = times9(M);
                M = c ? fp_add(m9, N) : fp_sub(m9, N);
            }
            v = fp_mul(A, M);
        }
        const Fq pv = shfl_xor(v, 1);
        Fq r = part ? fp_add(v, pv) : (odd ? fp_sub(pv, v) : fp_sub(v, pv));
        r = fp_add(r, shfl_xor(r, 2));
        r = fp_add(r, shfl_xor(r, 4));
        r = fp_add(r, shfl_xor(r, 8));
        if (s == 0) st(&dst[o], r);
    }
    __syncthreads();
}

// dst = a^(p^k), k = 1..3: word (i, part) = conj^k(a_i) * GAMMA_k[i]; lanes 0 and 1 of the word's group hold the two products
static __device__ __noinline__ void op_frobenius(Fq* dst, const Fq* a, int k, int t) {
    const Fq2* g = (k == 1) ? GAMMA1 : (k == 2) ? GAMMA2 : GAMMA3;
    const int o = t >> 4, s = t & 15, i = o >> 1, part = o & 1, odd = s & 1;
    const bool act = t < NPROD && s < 2;
    Fq A = fp_zero<FQ>();
    if (act) A = ld(&a[2 * i + odd]);
    __syncthreads();
    if (t < NPROD) {
        Fq v = fp_zero<FQ>();
        if (act) {
            const bool second = (part ^ odd) != 0;   // re: a_re g0, a_im g1; im: a_re g1, a_im g0
            v = fp_mul(A, second ? g[i].c1 : g[i].c0);
        }
        const Fq pv = shfl_xor(v, 1);
        const bool sub = ((part ^ (k & 1)) == 0);
        const Fq r = sub ? (odd ? fp_sub(pv, v) : fp_sub(v, pv)) : fp_add(v, pv);
        if (s == 0) st(&dst[o], r);
    }
    __syncthreads();
}
// a^(p^6): negate the odd powers of w.  Every thread reads and writes the same word: dst may alias a.
__device__ __forceinline__ void op_conj(Fq* dst, const Fq* a, int t) {
    if (t < 12) {
        const Fq v = ld(&a[t]);
        st(&dst[t], ((t >> 1) & 1) ? fp_neg(v) : v);
    }
    __syncthreads();
}
__device__ __forceinline__ void op_copy(Fq* dst, const Fq* a, int t) {
    if (t < 12) st(&dst[t], ld(&a[t]));
    __syncthreads();
}
// dst = xi * a word by word (dst must not alias a)
__device__ __forceinline__ void op_xi(Fq* dst, const Fq* a, int t) {
    if (t < 12) {
        const Fq m9 = times9(ld(&a[t])), n = ld(&a[t ^ 1]);
        st(&dst[t], (t & 1) ? fp_add(m9, n) : fp_sub(m9, n));
    }
    __syncthreads();
}

// dst = a^u, a in the cyclotomic subgroup; signed digits of u, MSB first
static __device__ __noinline__ void op_exp_by_u(Smem& sm, Fq* dst, const Fq* a, int t) {
    op_copy(sm.reg[BASE], a, t);
    op_xi(sm.reg[BASEX], a, t);
    op_conj(sm.reg[BASEC], sm.reg[BASE], t);
    op_conj(sm.reg[BASECX], sm.reg[BASEX], t);
    op_copy(dst, a, t);
#pragma unroll 1
    for (int d = 1; d < SNARKV_U_NAF_LEN; ++d) {
        op_mul(dst, dst, dst, nullptr, t);
        const int digit = U_NAF[d];
        if (digit > 0) op_mul(dst, dst, sm.reg[BASE], sm.reg[BASEX], t);
        else if (digit < 0) op_mul(dst, dst, sm.reg[BASEC], sm.reg[BASECX], t);
    }
}

// tower slot m (c0.c0.c0, c0.c0.c1, c0.c1.c0, ... serialisation order) -> w-basis word
__device__ __forceinline__ int tower_to_w(int m) {
    const int slot = m >> 1, part = m & 1;
    const int wp = (slot < 3) ? 2 * slot : 2 * (slot - 3) + 1;
    return 2 * wp + part;
}

}  // namespace fast

using namespace fast;

// Once per deciding key: the Fq2 constants of the merged two-pair line of every Miller step, as rows of 32 Fq words
// (lane = 2 slot + part), for the three cases both pairs live / only pair 0 / only pair 1.
//   both:  slot 0 cy1 cy2 | 1 xi c01 c02 | 2 cy1 cx2 | 3 cx1 cy2 | 4 cx1 cx2 | 6 cy1 c02 | 7 c01 cy2 | 8 cx1 c02 | 9 c01 cx2
//   only z: slot 0 cy_z | 2 cx_z | 6 c0_z
// (l_z(P) = cy_z y + cx_z x w + c0_z w^3; slots 2m, 2m + 1 add up to the coefficient of w^m of l_1 l_2.)
__global__ void __launch_bounds__(NSLOT) k_pair_tables(const uint8_t* __restrict__ coeffs, uint8_t* __restrict__ tables) {
    const int step = blockIdx.x, slot = threadIdx.x;
    const uint8_t* c1 = coeffs + (size_t)step * 192;
    const uint8_t* c2 = coeffs + (size_t)(NSTEP + step) * 192;
    const Fq2 cy1 = fq2_load(c1), cx1 = fq2_load(c1 + 64), c01 = fq2_load(c1 + 128);
    const Fq2 cy2 = fq2_load(c2), cx2 = fq2_load(c2 + 64), c02 = fq2_load(c2 + 128);
    Fq2 both = fq2_zero(), only0 = fq2_zero(), only1 = fq2_zero();
    switch (slot) {
        case 0: both = fq2_mul(cy1, cy2); only0 = cy1; only1 = cy2; break;
        case 1: both = fq2_mul_xi(fq2_mul(c01, c02)); break;
        case 2: both = fq2_mul(cy1, cx2); only0 = cx1; only1 = cx2; break;
        case 3: both = fq2_mul(cx1, cy2); break;
        case 4: both = fq2_mul(cx1, cx2); break;
        case 6: both = fq2_mul(cy1, c02); only0 = c01; only1 = c02; break;
        case 7: both = fq2_mul(c01, cy2); break;
        case 8: both = fq2_mul(cx1, c02); break;
        case 9: both = fq2_mul(c01, cx2); break;
        default: break;
    }
    const size_t row = (size_t)step * 32 + 2 * slot;
    fq2_store(tables + ((size_t)0 * NSTEP * 32 + row) * 32, both);
    fq2_store(tables + ((size_t)1 * NSTEP * 32 + row) * 32, only0);
    fq2_store(tables + ((size_t)2 * NSTEP * 32 + row) * 32, only1);
}

__global__ void __launch_bounds__(NT, 2) k_kzg_decide_fast(const uint8_t* __restrict__ lhs, const uint8_t* __restrict__ rhs, size_t N, int format,
                                                           const uint8_t* __restrict__ tables, const int* __restrict__ infinity,
                                                           uint8_t* __restrict__ accept, uint8_t* __restrict__ gt_out) {
    extern __shared__ __align__(16) uint8_t fast_smem[];
    Smem& sm = *reinterpret_cast<Smem*>(fast_smem);
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
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

        if (live0 || live1) {
            // ---- per-check scalars, then all merged lines (one warp per step) ---------------------------------------------
            const int mode = (live0 && live1) ? 0 : (live0 ? 1 : 2);
            if (t < NSLOT) {
                Fq v = fp_zero<FQ>();
                if (mode == 0) {
                    const Fq x1 = ld(&sm.px[0]), y1 = ld(&sm.py[0]), x2 = ld(&sm.px[1]), y2 = ld(&sm.py[1]);
                    switch (t) {
                        case 0: v = fp_mul(y1, y2); break;
                        case 1: v = fp_one<FQ>(); break;
                        case 2: v = fp_mul(y1, x2); break;
                        case 3: v = fp_mul(x1, y2); break;
                        case 4: v = fp_mul(x1, x2); break;
                        case 6: v = y1; break;
                        case 7: v = y2; break;
                        case 8: v = x1; break;
                        case 9: v = x2; break;
                        default: break;
                    }
                } else {
                    const int z = mode - 1;
                    if (t == 0) v = ld(&sm.py[z]);
                    else if (t == 2) v = ld(&sm.px[z]);
                    else if (t == 6) v = fp_one<FQ>();
                }
                st(&sm.S[t], v);
            }
            __syncthreads();
            {
                const uint8_t* tab = tables + (size_t)mode * NSTEP * 32 * 32;
                const int slot = lane >> 1, part = lane & 1;
                const Fq sc = ld(&sm.S[slot]);
#pragma unroll 1
                for (int step = wid; step < NSTEP; step += NT / 32) {
                    const Fq kq = fp_load<FQ>(tab + ((size_t)step * 32 + lane) * 32);
                    const Fq v = fp_mul(kq, sc);
                    const Fq r = fp_add(v, shfl_xor(v, 2));
                    const Fq other = shfl_xor(r, 1);
                    const Fq m9 = times9(r);
                    const Fq x = part ? fp_add(m9, other) : fp_sub(m9, other);
                    if ((slot & 1) == 0 && slot < 12) {
                        st(&sm.line[step][0][2 * (slot >> 1) + part], r);
                        st(&sm.line[step][1][2 * (slot >> 1) + part], x);
                    }
                }
            }
            __syncthreads();

            // ---- multi-Miller loop: f <- f^2, f <- f * (l_1 l_2) ------------------------------------------------------------
            int idx = 0;
#pragma unroll 1
            for (int b = NAF_LEN - 2; b >= 0; --b) {
                if (b != NAF_LEN - 2) op_mul(f, f, f, nullptr, t);
                op_mul(f, f, sm.line[idx][0], sm.line[idx][1], t);
                ++idx;
                if (ATE_NAF[b] != 0) {
                    op_mul(f, f, sm.line[idx][0], sm.line[idx][1], t);
                    ++idx;
                }
            }
            op_mul(f, f, sm.line[idx][0], sm.line[idx][1], t);
            op_mul(f, f, sm.line[idx + 1][0], sm.line[idx + 1][1], t);
        }

        // ---- final exponentiation -----------------------------------------------------------------------------------------
        // f^-1 by norms: c = conj(f), t = f c in Fq6, s = t^(p^2) t^(p^4), n = t s in Fq2, f^-1 = c s / n
        Fq* t0 = sm.reg[T0];
        Fq* t1 = sm.reg[T1];
        Fq* t2 = sm.reg[T2];
        op_conj(t1, f, t);
        op_mul(t0, f, t1, nullptr, t);
        op_frobenius(t2, t0, 2, t);
        op_frobenius(sm.reg[Y0], t2, 2, t);
        op_mul(t2, t2, sm.reg[Y0], nullptr, t);
        op_mul(sm.reg[Y0], t0, t2, nullptr, t);
        if (t == 0) {
            const Fq n0 = ld(&sm.reg[Y0][0]), n1 = ld(&sm.reg[Y0][1]);
            const Fq dinv = fp_inv_serial(fp_add(fp_mul(n0, n0), fp_mul(n1, n1)));
            st(&sm.reg[Y0][0], fp_mul(n0, dinv));
            st(&sm.reg[Y0][1], fp_neg(fp_mul(n1, dinv)));
        } else if (t < 12 && t >= 2) {
            st(&sm.reg[Y0][t], fp_zero<FQ>());
        }
        __syncthreads();
        op_mul(t2, t2, sm.reg[Y0], nullptr, t);
        op_mul(t0, t1, t2, nullptr, t);                              // f^-1
        // easy part: f <- conj(f) * f^-1 ; f <- f^(p^2) * f
        op_mul(f, t1, t0, nullptr, t);
        op_frobenius(t0, f, 2, t);
        op_mul(f, t0, f, nullptr, t);
        // hard part (p^4 - p^2 + 1)/r: y0 y1^2 y2^6 y3^12 y4^18 y5^30 y6^36 (Devegili-Scott-Dahab)
        op_exp_by_u(sm, sm.reg[FU], f, t);
        op_exp_by_u(sm, sm.reg[FU2], sm.reg[FU], t);
        op_exp_by_u(sm, sm.reg[FU3], sm.reg[FU2], t);
        op_frobenius(sm.reg[Y0], f, 1, t);                           // y0 = f^p f^(p^2) f^(p^3)
        op_frobenius(t0, f, 2, t);
        op_mul(sm.reg[Y0], sm.reg[Y0], t0, nullptr, t);
        op_frobenius(t0, f, 3, t);
        op_mul(sm.reg[Y0], sm.reg[Y0], t0, nullptr, t);
        op_conj(sm.reg[Y1], f, t);                                   // y1 = 1/f
        op_frobenius(sm.reg[Y2], sm.reg[FU2], 2, t);                 // y2 = (f^(u^2))^(p^2)
        op_frobenius(t0, sm.reg[FU], 1, t);                          // y3 = 1/(f^u)^p
        op_conj(sm.reg[Y3], t0, t);
        op_frobenius(t0, sm.reg[FU2], 1, t);                         // y4 = 1/(f^u (f^(u^2))^p)
        op_mul(t0, t0, sm.reg[FU], nullptr, t);
        op_conj(sm.reg[Y4], t0, t);
        op_conj(sm.reg[Y5], sm.reg[FU2], t);                         // y5 = 1/f^(u^2)
        op_frobenius(t0, sm.reg[FU3], 1, t);                         // y6 = 1/(f^(u^3) (f^(u^3))^p)
        op_mul(t0, t0, sm.reg[FU3], nullptr, t);
        op_conj(sm.reg[Y6], t0, t);
        op_mul(t0, sm.reg[Y6], sm.reg[Y6], nullptr, t);              // t0 = y6^2 y4 y5
        op_mul(t0, t0, sm.reg[Y4], nullptr, t);
        op_mul(t0, t0, sm.reg[Y5], nullptr, t);
        op_mul(t1, sm.reg[Y3], sm.reg[Y5], nullptr, t);              // t1 = y3 y5 t0
        op_mul(t1, t1, t0, nullptr, t);
        op_mul(t0, t0, sm.reg[Y2], nullptr, t);                      // t0 = t0 y2
        op_mul(t1, t1, t1, nullptr, t);                              // t1 = (t1^2 t0)^2
        op_mul(t1, t1, t0, nullptr, t);
        op_mul(t1, t1, t1, nullptr, t);
        op_mul(t0, t1, sm.reg[Y1], nullptr, t);                      // t0 = t1 y1
        op_mul(t1, t1, sm.reg[Y0], nullptr, t);                      // t1 = t1 y0
        op_mul(t0, t0, t0, nullptr, t);                              // gt = t0^2 t1
        op_mul(f, t0, t1, nullptr, t);

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

// called by kzg_set_key after k_g2_prepare: (re)builds the merged-line tables of the key on the stream
int kzg_build_pair_tables(snarkv_ctx* ctx) {
    const size_t bytes = (size_t)3 * NSTEP * 32 * 32;
    if (!ctx->d_key_tables) SNARKV_CUDA_TRY(ctx, cudaMalloc(&ctx->d_key_tables, bytes));
    k_pair_tables<<<NSTEP, NSLOT, 0, ctx->stream>>>((const uint8_t*)ctx->d_key_coeffs, (uint8_t*)ctx->d_key_tables);
    SNARKV_LAUNCH_CHECK(ctx, "k_pair_tables");
    ctx->launches++;
    return SNARKV_OK;
}

int kzg_decide_fast_device(snarkv_ctx* ctx, const void* d_lhs, const void* d_rhs, size_t N, int format, void* d_accept, void* d_gt) {
    const uint8_t* base = (const uint8_t*)ctx->d_key_coeffs;
    const int* d_inf = (const int*)(base + (size_t)2 * NSTEP * 192);
    const size_t smem = sizeof(Smem);
    if (ctx->pf_blocks_per_sm == 0) {
        SNARKV_CUDA_TRY(ctx, cudaFuncSetAttribute(k_kzg_decide_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0;
        SNARKV_CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_kzg_decide_fast, NT, smem));
        ctx->pf_blocks_per_sm = per_sm > 0 ? per_sm : 1;
    }
    const size_t cap = (size_t)ctx->sm_count * ctx->pf_blocks_per_sm;
    const unsigned blocks = (unsigned)(N < cap ? N : cap);
    Stage sg(ctx, "kzg_decide_fast");
    k_kzg_decide_fast<<<blocks, NT, smem, ctx->stream>>>((const uint8_t*)d_lhs, (const uint8_t*)d_rhs, N, format,
                                                         (const uint8_t*)ctx->d_key_tables, d_inf, (uint8_t*)d_accept, (uint8_t*)d_gt);
    SNARKV_LAUNCH_CHECK(ctx, "k_kzg_decide_fast");
    sg.launched();
    return SNARKV_OK;
}

}  // namespace snarkv
