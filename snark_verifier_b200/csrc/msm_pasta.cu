// msm_pasta.cu — the MSM pipeline of msm.cu compiled a second time over the PALLAS curve (y^2 = x^3 + 5 over the 255-bit Pallas base
// field), plus the IPA decider that consumes it.
//
// Replaces, for C = pasta::pallas::Affine (paths relative to snark-verifier/src):
//   util/msm.rs:259-343   multi_scalar_multiplication (Pippenger; chunk-parallel)  — the reference's only in-tree caller of a 2^k-term MSM
//   pcs/ipa.rs:401-417    h_coeffs: coeffs[j] = scalar * prod_{bit i of j set} xi[k - 1 - i]
//   pcs/ipa/decider.rs:47-70   IpaAs::decide: accept iff u == msm(h_coeffs(xi, 1), dk.g).to_affine(); decide_all loops it
//
// How: every kernel of the pipeline is written against `Fq` / `Fr` / SNARKV_CURVE_B / SNARKV_FIELD_BITS (fp.cuh).  This unit
// defines SNARKV_CURVE_PALLAS, which selects the generated Pallas field programs (fp_ptx_pallas.inc; gen_field_ptx.py --curve=pallas,
// emulator-verified for the 255-bit moduli), and renames the namespace so that both builds link into one library.  Differences
// from the BN254 build: no GLV split (glv.cuh holds BN254's lattice), ceil(256 / c) windows, generator (-1, 2).
#define SNARKV_CURVE_PALLAS 1
#define snarkv snarkv_pallas
#include "msm.cu"

namespace snarkv {

// pcs/ipa.rs:401-417: one thread per coefficient; bit i of the index selects xi[k - 1 - i] (`xi.iter().rev()` doubles the vector)
__global__ void __launch_bounds__(256) k_ipa_h_coeffs(const uint8_t* __restrict__ xi, uint32_t k, const uint8_t* __restrict__ scalar, int format,
                                                      uint8_t* __restrict__ out) {
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ((size_t)1 << k)) return;
    Fr acc = fp_load<FR>(scalar);
    if (format == SNARKV_CANONICAL) acc = fp_to_mont(acc);
    for (uint32_t i = 0; i < k; ++i) {
        if ((j >> i) & 1u) {
            Fr x = fp_load<FR>(xi + (size_t)(k - 1 - i) * 32);
            if (format == SNARKV_CANONICAL) x = fp_to_mont(x);
            acc = fp_mul(acc, x);
        }
    }
    fp_store<FR>(out + j * 32, acc);   // Montgomery form: what the pipeline reads with scalar_format = SNARKV_MONTGOMERY
}

// accept[0] = (u == result), both affine points in `format`; (0, 0) is the identity on both sides
__global__ void k_ipa_compare(const uint8_t* __restrict__ u, const uint8_t* __restrict__ result, uint8_t* __restrict__ accept) {
    const G1Affine a = g1_affine_load(u, 0), b = g1_affine_load(result, 0);
    accept[0] = (fp_eq(a.x, b.x) && fp_eq(a.y, b.y)) ? 1 : 0;
}

int ipa_h_coeffs_device(snarkv_ctx* ctx, const void* d_xi, uint32_t k, const void* d_scalar, int format, void* d_out_mont) {
    Stage sg(ctx, "ipa_h_coeffs");
    const size_t n = (size_t)1 << k;
    k_ipa_h_coeffs<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>((const uint8_t*)d_xi, k, (const uint8_t*)d_scalar, format, (uint8_t*)d_out_mont);
    SNARKV_LAUNCH_CHECK(ctx, "k_ipa_h_coeffs");
    sg.launched();
    return SNARKV_OK;
}

}  // namespace snarkv

using namespace snarkv;

#define PCTX_GUARD(ctx)                                                                         \
    do {                                                                                        \
        if (!(ctx)) return SNARKV_ERR_USAGE;                                                    \
        (ctx)->err.clear();                                                                     \
        cudaError_t _g = cudaSetDevice((ctx)->device);                                          \
        if (_g != cudaSuccess) return (ctx)->fail(SNARKV_ERR_CUDA, "cudaSetDevice", _g);        \
    } while (0)
static int pasta_bad_format(int f) { return f != SNARKV_CANONICAL && f != SNARKV_MONTGOMERY; }

extern "C" {

int snarkv_pallas_msm(snarkv_ctx* ctx, const uint8_t* scalars, const uint8_t* points, size_t n, int format, int flags, uint8_t out_affine[64]) {
    PCTX_GUARD(ctx);
    if (!scalars || !points || !out_affine || pasta_bad_format(format)) return ctx->fail(SNARKV_ERR_USAGE, "snarkv_pallas_msm: bad argument");
    ctx->profile_begin_call();
    return msm_run_host(ctx, scalars, points, n, format, flags, out_affine, nullptr);
}

int snarkv_pallas_msm_device(snarkv_ctx* ctx, const void* d_scalars, const void* d_points, size_t n, int format, int flags, void* d_out_affine,
                             void* d_out_jacobian, void* d_status) {
    PCTX_GUARD(ctx);
    if (!d_scalars || !d_points || (!d_out_affine && !d_out_jacobian) || pasta_bad_format(format))
        return ctx->fail(SNARKV_ERR_USAGE, "snarkv_pallas_msm_device: bad argument");
    if (n == 0) return ctx->fail(SNARKV_ERR_EMPTY, "multi_scalar_multiplication on an empty slice");
    ctx->profile_begin_call();
    if ((flags & SNARKV_CHECK_INPUTS) && !d_status) return ctx->fail(SNARKV_ERR_USAGE, "SNARKV_CHECK_INPUTS needs a status word");
    return msm_run_device(ctx, d_scalars, d_points, n, format, format, format, flags, d_out_affine, d_out_jacobian, d_status);
}

int snarkv_pallas_synth_scalars_device(snarkv_ctx* ctx, uint64_t seed, uint64_t start, size_t n, int format, void* d_out) {
    PCTX_GUARD(ctx);
    if (!d_out || pasta_bad_format(format)) return ctx->fail(SNARKV_ERR_USAGE, "snarkv_pallas_synth_scalars_device: bad argument");
    return synth_scalars_device(ctx, seed, start, n, format, d_out);
}
int snarkv_pallas_synth_points_device(snarkv_ctx* ctx, uint64_t seed, uint64_t start, size_t n, int format, void* d_out) {
    PCTX_GUARD(ctx);
    if (!d_out || pasta_bad_format(format)) return ctx->fail(SNARKV_ERR_USAGE, "snarkv_pallas_synth_points_device: bad argument");
    return synth_points_device(ctx, seed, start, n, format, d_out);
}

int snarkv_pallas_debug_field_op(snarkv_ctx* ctx, int field, int op, const uint8_t* a, const uint8_t* b, size_t n, uint8_t* out) {
    PCTX_GUARD(ctx);
    if (!a || !b || !out || n == 0 || field < 0 || field > 1) return ctx->fail(SNARKV_ERR_USAGE, "snarkv_pallas_debug_field_op: bad argument");
    uint8_t* d_a = (uint8_t*)ctx->wsget(WS_IO_A, n * 32);
    uint8_t* d_b = (uint8_t*)ctx->wsget(WS_IO_B, n * 32);
    uint8_t* d_o = (uint8_t*)ctx->wsget(WS_IO_C, n * 32);
    if (!d_a || !d_b || !d_o) return SNARKV_ERR_CUDA;
    cudaStream_t st = ctx->stream;
    SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(d_a, a, n * 32, cudaMemcpyHostToDevice, st));
    SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(d_b, b, n * 32, cudaMemcpyHostToDevice, st));
    int rc = field_op_device(ctx, field, op, d_a, d_b, n, d_o);
    if (rc) return rc;
    SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(out, d_o, n * 32, cudaMemcpyDeviceToHost, st));
    SNARKV_CUDA_TRY(ctx, cudaStreamSynchronize(st));
    return SNARKV_OK;
}

int snarkv_pallas_h_coeffs(snarkv_ctx* ctx, const uint8_t* xi, size_t k, const uint8_t scalar[32], int format, uint8_t* out) {
    PCTX_GUARD(ctx);
    if (!xi || !scalar || !out || k == 0 || k > 28 || pasta_bad_format(format)) return ctx->fail(SNARKV_ERR_USAGE, "snarkv_pallas_h_coeffs: bad argument");
    const size_t n = (size_t)1 << k;
    uint8_t* d_in = (uint8_t*)ctx->wsget(WS_IO_C, (k + 1) * 32);
    uint8_t* d_h = (uint8_t*)ctx->wsget(WS_IO_A, n * 32);
    if (!d_in || !d_h) return SNARKV_ERR_CUDA;
    cudaStream_t st = ctx->stream;
    SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(d_in, xi, k * 32, cudaMemcpyHostToDevice, st));
    SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(d_in + k * 32, scalar, 32, cudaMemcpyHostToDevice, st));
    int rc = ipa_h_coeffs_device(ctx, d_in, (uint32_t)k, d_in + k * 32, format, d_h);
    if (rc) return rc;
    if (format == SNARKV_CANONICAL && (rc = fr_from_mont_device(ctx, d_h, n)) != 0) return rc;
    SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(out, d_h, n * 32, cudaMemcpyDeviceToHost, st));
    SNARKV_CUDA_TRY(ctx, cudaStreamSynchronize(st));
    return SNARKV_OK;
}

int snarkv_ipa_set_deciding_key(snarkv_ctx* ctx, const uint8_t* g, size_t n, int format, int flags) {
    PCTX_GUARD(ctx);
    if (!g || n == 0 || (n & (n - 1)) != 0 || pasta_bad_format(format)) return ctx->fail(SNARKV_ERR_USAGE, "snarkv_ipa_set_deciding_key: g must hold 2^k points");
    if (ctx->d_ipa_g) { cudaStreamSynchronize(ctx->stream); cudaFree(ctx->d_ipa_g); ctx->d_ipa_g = nullptr; ctx->ipa_n = 0; }
    uint8_t* d_in = (uint8_t*)ctx->wsget(WS_IO_B, n * 64);
    uint8_t* d_o = (uint8_t*)ctx->wsget(WS_OUT, 1024);
    if (!d_in || !d_o) return SNARKV_ERR_CUDA;
    void* d_g = nullptr;
    SNARKV_CUDA_TRY(ctx, cudaMalloc(&d_g, 2 * n * 64));   // msm_bases_prepare writes [P | endomorphism slot] like the BN254 resident bases
    cudaStream_t st = ctx->stream;
    cudaError_t ce = cudaMemcpyAsync(d_in, g, n * 64, cudaMemcpyHostToDevice, st);
    int rc = ce == cudaSuccess ? msm_bases_prepare(ctx, d_in, n, format, (flags & SNARKV_CHECK_INPUTS) ? 1 : 0, d_g, d_o + 512) : SNARKV_ERR_CUDA;
    int status = 0;
    if (rc == SNARKV_OK) ce = cudaMemcpyAsync(&status, d_o + 512, 4, cudaMemcpyDeviceToHost, st);
    if (rc == SNARKV_OK && ce == cudaSuccess) ce = cudaStreamSynchronize(st);
    if (rc != SNARKV_OK || ce != cudaSuccess || status != 0) {
        cudaFree(d_g);
        if (rc != SNARKV_OK) return rc;
        if (ce != cudaSuccess) return ctx->fail(SNARKV_ERR_CUDA, "snarkv_ipa_set_deciding_key", ce);
        return ctx->fail(status, "committing key holds a point that is not a valid Pallas point");
    }
    ctx->d_ipa_g = d_g;
    ctx->ipa_n = n;
    return SNARKV_OK;
}

int snarkv_ipa_decide_batch(snarkv_ctx* ctx, const uint8_t* u, const uint8_t* xi, size_t k, size_t N, int format, uint8_t* accept) {
    PCTX_GUARD(ctx);
    if (!ctx->d_ipa_g) return ctx->fail(SNARKV_ERR_NO_KEY, "snarkv_ipa_decide_batch: no deciding key installed");
    if (N == 0) return SNARKV_OK;
    if (!u || !xi || !accept) return ctx->fail(SNARKV_ERR_USAGE, "snarkv_ipa_decide_batch: bad argument");
    if (format != SNARKV_CANONICAL) return ctx->fail(SNARKV_ERR_USAGE, "snarkv_ipa_decide_batch: SNARKV_CANONICAL only (accumulators come off a transcript)");
    ctx->profile_begin_call();
    if (((size_t)1 << k) != ctx->ipa_n) return ctx->fail(SNARKV_ERR_USAGE, "snarkv_ipa_decide_batch: xi must hold log2(|g|) challenges per accumulator");
    const size_t n = ctx->ipa_n;
    uint8_t* d_in = (uint8_t*)ctx->wsget(WS_IO_C, N * (k * 32 + 64) + 64);
    uint8_t* d_h = (uint8_t*)ctx->wsget(WS_IO_A, n * 32);
    uint8_t* d_o = (uint8_t*)ctx->wsget(WS_OUT, 1024);
    uint8_t* d_acc = (uint8_t*)ctx->wsget(WS_FR_PROG, N + 64);
    if (!d_in || !d_h || !d_o || !d_acc) return SNARKV_ERR_CUDA;
    cudaStream_t st = ctx->stream;
    uint8_t* d_xi = d_in;
    uint8_t* d_u = d_in + N * k * 32;
    uint8_t* d_one = d_u + N * 64;
    uint8_t one[32] = {1};
    SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(d_xi, xi, N * k * 32, cudaMemcpyHostToDevice, st));
    SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(d_u, u, N * 64, cudaMemcpyHostToDevice, st));
    SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(d_one, one, 32, cudaMemcpyHostToDevice, st));
    for (size_t a = 0; a < N; ++a) {
        int rc = ipa_h_coeffs_device(ctx, d_xi + a * k * 32, (uint32_t)k, d_one, SNARKV_CANONICAL, d_h);   // h_coeffs(&xi, C::Scalar::ONE)
        if (rc) return rc;
        // the resident committing key is already Montgomery and validated; the result comes back canonical, like u
        rc = msm_run_device(ctx, d_h, ctx->d_ipa_g, n, SNARKV_MONTGOMERY, SNARKV_MONTGOMERY, SNARKV_CANONICAL, 0, d_o, nullptr, d_o + 512);
        if (rc) return rc;
        k_ipa_compare<<<1, 1, 0, st>>>(d_u + a * 64, d_o, d_acc + a);
        SNARKV_LAUNCH_CHECK(ctx, "k_ipa_compare");
        ctx->launches++;
    }
    SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(accept, d_acc, N, cudaMemcpyDeviceToHost, st));
    SNARKV_CUDA_TRY(ctx, cudaStreamSynchronize(st));
    return SNARKV_OK;
}

}  // extern "C"
