// pairing.cu — batched KZG accumulator decision on sm_100a: optimal-ate multi-Miller loop + final exponentiation.
//
// Replaces, behind `snarkv_kzg_set_deciding_key` / `snarkv_kzg_decide_batch*` (include/snarkv_cuda.h), the reference's
//   KzgDecidingKey::new                                  snark-verifier/src/pcs/kzg/decider.rs:6-42
//   <KzgAs as AccumulationDecider<_, NativeLoader>>::decide      decider.rs:70-82
//   ...::decide_all                                              decider.rs:84-93
// and the halo2curves 0.6.0 calls they make (decider.rs:74-78): `G2Prepared::from`, `Bn256::multi_miller_loop`,
// `MillerLoopResult::final_exponentiation`, `Gt::is_identity`.
//
// Differences in shape, not in result:
//   * the reference rebuilds G2Prepared for g2 and -s_g2 inside every decide call (decider.rs:74); here the line
//     coefficients of both fixed G2 points are computed ONCE per deciding key by k_g2_prepare and kept in HBM
//     (2 x 88 x 192 B = 33 KB, read uniformly by every thread so they stay in L1/L2);
//   * decide_all loops one check at a time; here N checks run as N threads (one check per thread, Fq12 state in
//     thread-private memory), in two kernels: Miller loop (f -> HBM, 384 B/check) and final exponentiation.
// Per-check algorithmic bytes: 128 B in (lhs, rhs) + 1 B out (+ 384 B when GT is requested).
#include "ctx.hpp"
#include "g1.cuh"
#include "tower.cuh"

namespace snarkv {

struct alignas(16) LineCoeff {
    Fq2 cy, cx, c0;  // l(P) = cy * yP + cx * xP * w + c0 * w^3
};
struct G2Jac { Fq2 x, y, z; };
struct G2Aff { Fq2 x, y; };

static constexpr int NUM_COEFFS = SNARKV_ATE_NUM_COEFFS;
static constexpr int NAF_LEN = SNARKV_ATE_NAF_LEN;

// tangent at T, scaled by the Fq2 factor 2YZ^3 (killed by the final exponentiation):
//   cy = 2YZ * Z^2, cx = -3X^2 Z^2, c0 = 3X^3 - 2Y^2 ;  T <- 2T (dbl-2009-l over Fq2)
static __device__ __noinline__ LineCoeff g2_doubling_step(G2Jac& t) {
    Fq2 a = fq2_sqr(t.x), b = fq2_sqr(t.y), c = fq2_sqr(b);
    Fq2 zz = fq2_sqr(t.z);
    Fq2 e = fq2_add(fq2_dbl(a), a);
    Fq2 z3 = fq2_dbl(fq2_mul(t.y, t.z));
    LineCoeff l;
    l.cy = fq2_mul(z3, zz);
    l.cx = fq2_neg(fq2_mul(e, zz));
    l.c0 = fq2_sub(fq2_mul(e, t.x), fq2_dbl(b));
    Fq2 d = fq2_dbl(fq2_sub(fq2_sub(fq2_sqr(fq2_add(t.x, b)), a), c));
    Fq2 f = fq2_sqr(e);
    Fq2 x3 = fq2_sub(f, fq2_dbl(d));
    t.y = fq2_sub(fq2_mul(e, fq2_sub(d, x3)), fq2_dbl(fq2_dbl(fq2_dbl(c))));
    t.x = x3;
    t.z = z3;
    return l;
}
// chord through T and affine Q, scaled by H*Z:  cy = HZ, cx = -R, c0 = R x2 - y2 HZ ;  T <- T + Q
static __device__ __noinline__ LineCoeff g2_addition_step(G2Jac& t, const G2Aff& q) {
    Fq2 zz = fq2_sqr(t.z);
    Fq2 h = fq2_sub(fq2_mul(q.x, zz), t.x);
    Fq2 r = fq2_sub(fq2_mul(fq2_mul(q.y, zz), t.z), t.y);
    Fq2 z3 = fq2_mul(t.z, h);
    LineCoeff l;
    l.cy = z3;
    l.cx = fq2_neg(r);
    l.c0 = fq2_sub(fq2_mul(r, q.x), fq2_mul(q.y, z3));
    Fq2 hh = fq2_sqr(h), hhh = fq2_mul(h, hh), v = fq2_mul(t.x, hh);
    Fq2 x3 = fq2_sub(fq2_sub(fq2_sqr(r), hhh), fq2_dbl(v));
    t.y = fq2_sub(fq2_mul(r, fq2_sub(v, x3)), fq2_mul(t.y, hhh));
    t.x = x3;
    t.z = z3;
    return l;
}

__device__ __forceinline__ void line_store(uint8_t* base, int idx, const LineCoeff& l) {
    uint8_t* p = base + (size_t)idx * 192;
    fq2_store(p, l.cy);
    fq2_store(p + 64, l.cx);
    fq2_store(p + 128, l.c0);
}

// `G2Prepared::from(q)` for q = g2 (thread 0) and q = -s_g2 (thread 1).  Inputs CANONICAL 128 B each.
// flags_out[pair] = 1 if the point is the identity (pair contributes 1), status = BAD_POINT when not on the twist.
__global__ void k_g2_prepare(const uint8_t* __restrict__ g2s /* 2 x 128 B */, uint8_t* __restrict__ coeffs, int* __restrict__ infinity,
                             int* __restrict__ status) {
    const int pair = threadIdx.x;
    if (pair >= 2) return;
    const uint8_t* in = g2s + pair * 128;
    G2Aff q;
    q.x = fq2_load(in);
    q.y = fq2_load(in + 64);
    if (!fp_is_canonical(q.x.c0) || !fp_is_canonical(q.x.c1) || !fp_is_canonical(q.y.c0) || !fp_is_canonical(q.y.c1)) {
        atomicCAS(status, 0, SNARKV_ERR_BAD_POINT);
        return;
    }
    const bool inf = fq2_is_zero(q.x) && fq2_is_zero(q.y);
    infinity[pair] = inf ? 1 : 0;
    if (inf) return;
    q.x = {fp_to_mont(q.x.c0), fp_to_mont(q.x.c1)};
    q.y = {fp_to_mont(q.y.c0), fp_to_mont(q.y.c1)};
    // on the twist  y^2 = x^3 + 3/xi ?
    if (!fq2_eq(fq2_sqr(q.y), fq2_add(fq2_mul(fq2_sqr(q.x), q.x), TWIST_B))) {
        atomicCAS(status, 0, SNARKV_ERR_BAD_POINT);
        return;
    }
    if (pair == 1) q.y = fq2_neg(q.y);  // (-dk.s_g2).into()   decider.rs:74
    uint8_t* out = coeffs + (size_t)pair * NUM_COEFFS * 192;
    G2Jac t = {q.x, q.y, fq2_one()};
    G2Aff nq = {q.x, fq2_neg(q.y)};
    int idx = 0;
    for (int i = NAF_LEN - 2; i >= 0; --i) {
        line_store(out, idx++, g2_doubling_step(t));
        if (ATE_NAF[i] == 1) line_store(out, idx++, g2_addition_step(t, q));
        else if (ATE_NAF[i] == -1) line_store(out, idx++, g2_addition_step(t, nq));
    }
    // Q1 = pi(Q), Q2 = -pi^2(Q) on the twist
    G2Aff q1 = {fq2_mul(fq2_conj(q.x), GAMMA1[2]), fq2_mul(fq2_conj(q.y), GAMMA1[3])};
    G2Aff q2 = {fq2_mul(q.x, GAMMA2[2]), fq2_neg(fq2_mul(q.y, GAMMA2[3]))};
    line_store(out, idx++, g2_addition_step(t, q1));
    line_store(out, idx++, g2_addition_step(t, q2));
}

__device__ __forceinline__ void ell(Fq12& f, const uint8_t* __restrict__ coeffs, int idx, const G1Affine& p) {
    const uint8_t* c = coeffs + (size_t)idx * 192;
    Fq2 cy = fq2_load(c), cx = fq2_load(c + 64), c0 = fq2_load(c + 128);
    f = fq12_mul_by_034(f, fq2_scale(cy, p.y), fq2_scale(cx, p.x), c0);
}

__device__ __forceinline__ G1Affine load_g1_checked(const uint8_t* base, size_t i, int format, bool& ok) {
    G1Affine p = g1_affine_load(base, i);
    if (format == SNARKV_CANONICAL) {
        if (!fp_is_canonical(p.x) || !fp_is_canonical(p.y)) ok = false;
        p.x = fp_to_mont(p.x);
        p.y = fp_to_mont(p.y);
    }
    if (!g1_affine_is_on_curve(p)) ok = false;  // `from_xy(..).unwrap()` discipline of accumulator.rs:75-78
    return p;
}

// `Bn256::multi_miller_loop(&[(&lhs, &g2_prepared), (&rhs, &neg_s_g2_prepared)])`: shared squarings, pairs with an
// identity on either side skipped.  One thread per accumulator; f written to HBM (tower order, Montgomery).
__global__ void __launch_bounds__(64) k_miller_loop(const uint8_t* __restrict__ lhs, const uint8_t* __restrict__ rhs, size_t N, int format,
                                                    const uint8_t* __restrict__ coeffs, const int* __restrict__ infinity,
                                                    Fq12* __restrict__ f_out, uint8_t* __restrict__ bad) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    bool ok = true;
    G1Affine p0 = load_g1_checked(lhs, i, format, ok);
    G1Affine p1 = load_g1_checked(rhs, i, format, ok);
    bad[i] = ok ? 0 : 1;
    const bool live0 = ok && !g1_affine_is_identity(p0) && !infinity[0];
    const bool live1 = ok && !g1_affine_is_identity(p1) && !infinity[1];
    const uint8_t* c0 = coeffs;
    const uint8_t* c1 = coeffs + (size_t)NUM_COEFFS * 192;
    Fq12 f = fq12_one();
    int idx = 0;
    for (int b = NAF_LEN - 2; b >= 0; --b) {
        if (b != NAF_LEN - 2) f = fq12_sqr(f);
        if (live0) ell(f, c0, idx, p0);
        if (live1) ell(f, c1, idx, p1);
        ++idx;
        if (ATE_NAF[b] != 0) {
            if (live0) ell(f, c0, idx, p0);
            if (live1) ell(f, c1, idx, p1);
            ++idx;
        }
    }
    for (int extra = 0; extra < 2; ++extra) {
        if (live0) ell(f, c0, idx, p0);
        if (live1) ell(f, c1, idx, p1);
        ++idx;
    }
    f_out[i] = f;
}

static __device__ __noinline__ Fq12 exp_by_u(const Fq12& f) {
    Fq12 r = f;
    const uint64_t u = SNARKV_BN_U;  // 63 bits, bit 62 leading
    for (int i = 61; i >= 0; --i) {
        r = fq12_cyclotomic_sqr(r);
        if ((u >> i) & 1ull) r = fq12_mul(r, f);
    }
    return r;
}

// `MillerLoopResult::final_exponentiation` + `Gt::is_identity`: f^((p^12-1)/r) == 1.
// Easy part (p^6-1)(p^2+1); hard part (p^4-p^2+1)/r by the Devegili-Scott-Dahab chain y0 y1^2 y2^6 y3^12 y4^18 y5^30 y6^36.
__global__ void __launch_bounds__(64) k_final_exp(const Fq12* __restrict__ f_in, size_t N, const uint8_t* __restrict__ bad,
                                                  uint8_t* __restrict__ accept, uint8_t* __restrict__ gt_out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    Fq12 f = f_in[i];
    {
        Fq12 t = fq12_mul(fq12_conj(f), fq12_inv(f));
        f = fq12_mul(fq12_frobenius(t, 2), t);
    }
    Fq12 fu = exp_by_u(f);
    Fq12 fu2 = exp_by_u(fu);
    Fq12 fu3 = exp_by_u(fu2);
    Fq12 y0 = fq12_mul(fq12_mul(fq12_frobenius(f, 1), fq12_frobenius(f, 2)), fq12_frobenius(f, 3));
    Fq12 y4 = fq12_conj(fq12_mul(fu, fq12_frobenius(fu2, 1)));
    Fq12 y6 = fq12_conj(fq12_mul(fu3, fq12_frobenius(fu3, 1)));
    Fq12 y5 = fq12_conj(fu2);
    Fq12 t0 = fq12_mul(fq12_mul(fq12_cyclotomic_sqr(y6), y4), y5);
    Fq12 y3 = fq12_conj(fq12_frobenius(fu, 1));
    Fq12 t1 = fq12_mul(fq12_mul(y3, y5), t0);
    t0 = fq12_mul(t0, fq12_frobenius(fu2, 2));  // y2
    t1 = fq12_cyclotomic_sqr(fq12_mul(fq12_cyclotomic_sqr(t1), t0));
    t0 = fq12_mul(t1, fq12_conj(f));  // y1
    t1 = fq12_mul(t1, y0);
    Fq12 gt = fq12_mul(fq12_cyclotomic_sqr(t0), t1);
    const bool is_bad = bad[i] != 0;
    accept[i] = (!is_bad && fq12_is_one(gt)) ? 1 : 0;
    if (gt_out) {
        const Fq* c = reinterpret_cast<const Fq*>(&gt);
        uint8_t* o = gt_out + i * 384;
#pragma unroll 1
        for (int k = 0; k < 12; ++k) fp_store<FQ>(o + 32 * k, is_bad ? fp_zero<FQ>() : fp_from_mont(c[k]));
    }
}

// ---------------------------------------------------------------------------------------------------------------------
void kzg_free_key(snarkv_ctx* ctx) {
    if (ctx->d_key_coeffs) cudaFree(ctx->d_key_coeffs);
    ctx->d_key_coeffs = nullptr;
    if (ctx->d_key_tables) cudaFree(ctx->d_key_tables);
    ctx->d_key_tables = nullptr;
    ctx->has_key = false;
}

int kzg_set_key(snarkv_ctx* ctx, const uint8_t g1[64], const uint8_t g2[128], const uint8_t s_g2[128]) {
    cudaStream_t st = ctx->stream;
    const size_t coeff_bytes = (size_t)2 * NUM_COEFFS * 192;
    if (!ctx->d_key_coeffs) {
        // [coeffs | infinity[2] | status | staging 256 B]
        SNARKV_CUDA_TRY(ctx, cudaMalloc(&ctx->d_key_coeffs, coeff_bytes + 16 + 256));
    }
    uint8_t* base = (uint8_t*)ctx->d_key_coeffs;
    int* d_inf = (int*)(base + coeff_bytes);
    int* d_status = d_inf + 2;
    uint8_t* d_stage = base + coeff_bytes + 16;
    ctx->has_key = false;
    SNARKV_CUDA_TRY(ctx, cudaMemsetAsync(d_inf, 0, 16, st));
    SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(d_stage, g2, 128, cudaMemcpyHostToDevice, st));
    SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(d_stage + 128, s_g2, 128, cudaMemcpyHostToDevice, st));
    k_g2_prepare<<<1, 32, 0, st>>>(d_stage, base, d_inf, d_status);
    SNARKV_LAUNCH_CHECK(ctx, "k_g2_prepare");
    ctx->launches++;
    {   // merged two-pair line constants for the latency kernel (pairing_fast.cu); rows of a key at infinity are never read
        const int rc = kzg_build_pair_tables(ctx);
        if (rc) return rc;
    }
    int status = 0;
    SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(&status, d_status, 4, cudaMemcpyDeviceToHost, st));
    SNARKV_CUDA_TRY(ctx, cudaStreamSynchronize(st));
    if (status != 0) return ctx->fail(SNARKV_ERR_BAD_POINT, "deciding key: g2 / s_g2 is not a valid G2Affine");
    memcpy(ctx->key_g1, g1, 64);
    ctx->key_num_coeffs = NUM_COEFFS;
    ctx->has_key = true;
    return SNARKV_OK;
}

int kzg_decide_device(snarkv_ctx* ctx, const void* d_lhs, const void* d_rhs, size_t N, int format, void* d_accept, void* d_gt) {
    // Small and medium batches are latency-bound on one thread per check (13 ms per check regardless of N): give each check
    // a whole thread block instead (pairing_coop.cu).  Very large batches fill the machine either way; keep the leaner kernel.
    // modes: 0 auto, 1 thread per check, 2 cooperative (block or warp per check by N), 3 block per check, 4 warp per check
    // measured crossover on B200 (profiles/r01_pairing_modes.txt): warp-per-check 395 K checks/s vs thread-per-check 13.9 ms flat
    // mode 5 / small N: the latency kernel (pairing_fast.cu) — one 256-thread block per check
    if (ctx->pairing_mode == 5 || (ctx->pairing_mode == 0 && N <= (size_t)ctx->sm_count * ctx->pf_max_per_sm))
        return kzg_decide_fast_device(ctx, d_lhs, d_rhs, N, format, d_accept, d_gt);
    const bool coop = ctx->pairing_mode >= 2 || (ctx->pairing_mode == 0 && N <= (size_t)ctx->sm_count * 37);
    if (coop) return kzg_decide_coop_device(ctx, d_lhs, d_rhs, N, format, d_accept, d_gt);
    Fq12* f = (Fq12*)ctx->wsget(WS_PAIR_A, N * sizeof(Fq12));
    uint8_t* bad = (uint8_t*)ctx->wsget(WS_PAIR_B, N);
    if (!f || !bad) return SNARKV_ERR_CUDA;
    const uint8_t* base = (const uint8_t*)ctx->d_key_coeffs;
    const int* d_inf = (const int*)(base + (size_t)2 * NUM_COEFFS * 192);
    const unsigned blocks = (unsigned)((N + 63) / 64);
    {
        Stage sg(ctx, "kzg_miller_loop");
        k_miller_loop<<<blocks, 64, 0, ctx->stream>>>((const uint8_t*)d_lhs, (const uint8_t*)d_rhs, N, format, base, d_inf, f, bad);
        SNARKV_LAUNCH_CHECK(ctx, "k_miller_loop");
        sg.launched();
    }
    {
        Stage sg(ctx, "kzg_final_exp");
        k_final_exp<<<blocks, 64, 0, ctx->stream>>>(f, N, bad, (uint8_t*)d_accept, (uint8_t*)d_gt);
        SNARKV_LAUNCH_CHECK(ctx, "k_final_exp");
        sg.launched();
    }
    return SNARKV_OK;
}

}  // namespace snarkv
