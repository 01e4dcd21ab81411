// capi.cu — the extern "C" surface declared in include/snarkv_cuda.h.
//
// This file is the drop-in boundary for the reference's Loader / Decider hot path: the host-buffer entry points are what
// a Rust `impl EcPointLoader<G1Affine> for CudaLoader` (mirroring snark-verifier/src/loader/native.rs:43-72) and
// `impl AccumulationDecider<G1Affine, CudaLoader> for KzgAs<Bn256, MOS>` (mirroring pcs/kzg/decider.rs:62-94) call; see
// INTEGRATION.md.  There is no CPU arithmetic here — only copies, launches and status decoding.
#include <new>

#include <cstdlib>

#include <cstring>
#include <vector>
#include "ctx.hpp"

using namespace snarkv;

#define CTX_GUARD(ctx)                            \
    do {                                          \
        if (!(ctx)) return SNARKV_ERR_USAGE;      \
        (ctx)->err.clear();                       \
        cudaError_t _g = cudaSetDevice((ctx)->device); \
        if (_g != cudaSuccess) return (ctx)->fail(SNARKV_ERR_CUDA, "cudaSetDevice", _g); \
    } while (0)

static int bad_format(int f) { return f != SNARKV_CANONICAL && f != SNARKV_MONTGOMERY; }

// Host-buffer body of snarkv_g1_msm_batch_rlc, shared with the multi-device entry (multi.cu): segments j = 0..m-1 are scaled by
// rho^(first_power + j); the result is returned as affine bytes (out_affine) and/or left as a Jacobian partial on the device.
namespace snarkv {
int msm_batch_rlc_host(snarkv_ctx* ctx, const uint8_t* scalars, const uint8_t* points, const uint64_t* offsets, size_t m, const uint8_t rho[32],
                       int format, int flags, uint64_t first_power, uint8_t* out_affine, void* d_out_jacobian) {
    if (!scalars || !points || !offsets || !rho || bad_format(format) || m == 0)
        return ctx->fail(SNARKV_ERR_USAGE, "snarkv_g1_msm_batch_rlc: bad argument");
    const uint64_t base = offsets[0];
    for (size_t j = 0; j < m; ++j)
        if (offsets[j + 1] <= offsets[j]) return ctx->fail(SNARKV_ERR_EMPTY, "empty or decreasing MSM segment");
    const size_t total = offsets[m] - base;
    ctx->profile_begin_call();
    uint8_t* d_s = (uint8_t*)ctx->wsget(WS_IO_A, total * 32);
    uint8_t* d_p = (uint8_t*)ctx->wsget(WS_IO_B, total * 64);
    uint64_t* d_off = (uint64_t*)ctx->wsget(WS_IO_C, (m + 1) * 8);
    uint8_t* d_scaled = (uint8_t*)ctx->wsget(WS_IO_D, total * 32);
    uint8_t* d_pw = (uint8_t*)ctx->wsget(WS_IO_E, m * 32 + 32);   // [rho | powers]
    uint8_t* d_o = (uint8_t*)ctx->wsget(WS_OUT, 1024);            // [affine 64 | status 2 x 4 @64]
    if (!d_s || !d_p || !d_off || !d_scaled || !d_pw || !d_o) return SNARKV_ERR_CUDA;
    cudaStream_t st = ctx->stream;
    std::vector<uint64_t> rebased;
    const uint64_t* off = offsets;
    if (base != 0) {   // a shard of a longer batch: offsets relative to the shard's first term
        rebased.resize(m + 1);
        for (size_t j = 0; j <= m; ++j) rebased[j] = offsets[j] - base;
        off = rebased.data();
    }
    SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(d_pw, rho, 32, cudaMemcpyHostToDevice, st));
    SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(d_s, scalars + base * 32, total * 32, cudaMemcpyHostToDevice, st));
    SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(d_off, off, (m + 1) * 8, cudaMemcpyHostToDevice, st));
    SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(d_p, points + base * 64, total * 64, cudaMemcpyHostToDevice, st));
    int rc = msm_batch_rlc_device(ctx, d_s, d_p, d_off, m, total, d_pw, format, flags, d_scaled, d_pw + 32, d_o, d_o + 64, first_power,
                                  d_out_jacobian);
    if (rc) { cudaStreamSynchronize(st); return rc; }   // `rebased` (pageable) must have left the host before it is destroyed
    uint8_t host[72];
    SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(host, d_o, 72, cudaMemcpyDeviceToHost, st));
    SNARKV_CUDA_TRY(ctx, cudaStreamSynchronize(st));
    int st0, st1;
    memcpy(&st0, host + 64, 4);
    memcpy(&st1, host + 68, 4);
    const int status = st0 ? st0 : st1;
    if (status != 0) return ctx->fail(status, status == SNARKV_ERR_BAD_SCALAR ? "scalar is not a canonical Fr" : "point is not a valid G1Affine");
    if (out_affine) memcpy(out_affine, host, 64);
    return SNARKV_OK;
}

// Child context for a second concurrent MSM: own stream, own (grow-only) workspace, same device and tuning knobs.
snarkv_ctx* ctx_aux(snarkv_ctx* ctx) {
    if (ctx->aux) {
        snarkv_ctx* a = ctx->aux;   // knobs may have been changed through the C ABI since the child was made
        a->window_bits = ctx->window_bits; a->glv_mode = ctx->glv_mode; a->accumulate_mode = ctx->accumulate_mode;
        return a;
    }
    snarkv_ctx* a = new (std::nothrow) snarkv_ctx();
    if (!a) { ctx->fail(SNARKV_ERR_CUDA, "child context allocation"); return nullptr; }
    a->device = ctx->device;
    a->sm_count = ctx->sm_count;
    cudaError_t ce = cudaStreamCreateWithFlags(&a->own_stream, cudaStreamNonBlocking);
    if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&ctx->fork_ev, cudaEventDisableTiming);
    if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&ctx->join_ev, cudaEventDisableTiming);
    if (ce != cudaSuccess) {
        if (a->own_stream) cudaStreamDestroy(a->own_stream);
        if (ctx->fork_ev) { cudaEventDestroy(ctx->fork_ev); ctx->fork_ev = nullptr; }
        delete a;
        ctx->fail(SNARKV_ERR_CUDA, "child context stream / events", ce);
        return nullptr;
    }
    a->stream = a->own_stream;
    a->window_bits = ctx->window_bits; a->glv_mode = ctx->glv_mode; a->accumulate_mode = ctx->accumulate_mode;
    a->sort_blocks_per_sm = ctx->sort_blocks_per_sm; a->sort_tile = ctx->sort_tile;
    a->ba_k = ctx->ba_k; a->ba_pairs_min = ctx->ba_pairs_min; a->ba_q = ctx->ba_q; a->ba_min_load = ctx->ba_min_load;
    a->bc_r = ctx->bc_r; a->bc_auto = ctx->bc_auto; a->bc_min_load = ctx->bc_min_load;
    ctx->aux = a;
    return a;
}
}  // namespace snarkv

extern "C" {

const char* snarkv_version(void) { return "snarkv-cuda 0.1 (sm_100a; BN254 G1 MSM + KZG decide)"; }

int snarkv_init(int device, snarkv_ctx** out) {
    if (!out) return SNARKV_ERR_USAGE;
    *out = nullptr;
    int count = 0;
    cudaError_t ce = cudaGetDeviceCount(&count);
    if (ce != cudaSuccess || device < 0 || device >= count) return SNARKV_ERR_CUDA;  // no GPU: fail loudly, no fallback
    if (cudaSetDevice(device) != cudaSuccess) return SNARKV_ERR_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return SNARKV_ERR_CUDA;
    if (prop.major != 10) return SNARKV_ERR_CUDA;  // kernels are built for sm_100a only
    snarkv_ctx* c = new (std::nothrow) snarkv_ctx();
    if (!c) return SNARKV_ERR_USAGE;
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete c;
        return SNARKV_ERR_CUDA;
    }
    for (cudaEvent_t& e : c->copy_done)
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) {
            delete c;
            return SNARKV_ERR_CUDA;
        }
    c->stream = c->own_stream;
    // developer tuning knobs of the batched-affine accumulation (bucket_affine.cuh); the defaults are the measured optimum
    auto env_int = [](const char* name, int lo, int hi, int dflt) {
        const char* v = getenv(name);
        if (!v || !*v) return dflt;
        const int x = atoi(v);
        return x < lo ? lo : (x > hi ? hi : x);
    };
    c->ba_k = env_int("SNARKV_BA_K", 1, 128, c->ba_k);
    c->ba_pairs_min = env_int("SNARKV_BA_PAIRS_MIN", 1, 1 << 20, c->ba_pairs_min);
    c->ba_q = env_int("SNARKV_BA_Q", 1, 4, c->ba_q);
    c->host_chunks = env_int("SNARKV_HOST_CHUNKS", 2, 7, c->host_chunks);
    c->sort_blocks_per_sm = env_int("SNARKV_SORT_BLOCKS", 0, 16, 0);
    c->sort_tile = env_int("SNARKV_SORT_TILE", 4096, 8192, c->sort_tile) == 8192 ? 8192 : 4096;
    c->host_chunk_min_log_n = env_int("SNARKV_HOST_CHUNK_MIN", 16, 30, c->host_chunk_min_log_n);
    c->host_chunks_small = env_int("SNARKV_HOST_CHUNKS_SMALL", 2, 7, c->host_chunks_small);
    c->host_chunk_ratio_pct = env_int("SNARKV_HOST_RATIO", 100, 400, c->host_chunk_ratio_pct);
    c->ba_min_load = env_int("SNARKV_BA_MIN_LOAD", 1, 1 << 20, c->ba_min_load);
    c->bc_r = env_int("SNARKV_BC_R", 8, 16, c->bc_r);
    if (c->bc_r != 8 && c->bc_r != 12) c->bc_r = 16;
    c->bc_auto = env_int("SNARKV_BC_AUTO", 0, 1, c->bc_auto);
    c->bc_min_load = env_int("SNARKV_BC_MIN_LOAD", 1, 1 << 20, c->bc_min_load);
    c->pf_max_per_sm = env_int("SNARKV_PAIRING_FAST_MAX", 0, 1 << 20, c->pf_max_per_sm);
    c->overlap_msms = env_int("SNARKV_OVERLAP_MSMS", 0, 1, c->overlap_msms);
    *out = c;
    return SNARKV_OK;
}

void snarkv_destroy(snarkv_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    kzg_free_key(ctx);
    if (ctx->aux) {
        for (int i = 0; i < WS_SLOTS; ++i)
            if (ctx->aux->ws[i]) cudaFree(ctx->aux->ws[i]);
        for (cudaEvent_t e : ctx->aux->event_pool) cudaEventDestroy(e);
        if (ctx->aux->own_stream) cudaStreamDestroy(ctx->aux->own_stream);
        delete ctx->aux;
        ctx->aux = nullptr;
    }
    if (ctx->fork_ev) cudaEventDestroy(ctx->fork_ev);
    if (ctx->join_ev) cudaEventDestroy(ctx->join_ev);
    if (ctx->d_ipa_g) cudaFree(ctx->d_ipa_g);
    for (int i = 0; i < WS_SLOTS; ++i)
        if (ctx->ws[i]) cudaFree(ctx->ws[i]);
    for (cudaEvent_t e : ctx->event_pool) cudaEventDestroy(e);
    for (cudaEvent_t e : ctx->copy_done)
        if (e) cudaEventDestroy(e);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

const char* snarkv_last_error(const snarkv_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int snarkv_set_stream(snarkv_ctx* ctx, void* cuda_stream) {
    CTX_GUARD(ctx);
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return SNARKV_OK;
}

int snarkv_set_window_bits(snarkv_ctx* ctx, int c) {
    if (!ctx || c < 0 || c > 22) return SNARKV_ERR_USAGE;
    ctx->window_bits = c;
    return SNARKV_OK;
}

int snarkv_set_glv_mode(snarkv_ctx* ctx, int mode) {
    if (!ctx || mode < 0 || mode > 2) return SNARKV_ERR_USAGE;
    ctx->glv_mode = mode;
    return SNARKV_OK;
}

int snarkv_set_pairing_mode(snarkv_ctx* ctx, int mode) {
    if (!ctx || mode < 0 || mode > 5) return SNARKV_ERR_USAGE;
    ctx->pairing_mode = mode;
    return SNARKV_OK;
}

int snarkv_set_accumulate_mode(snarkv_ctx* ctx, int mode) {
    if (!ctx || mode < 0 || mode > 5) return SNARKV_ERR_USAGE;
    ctx->accumulate_mode = mode;
    return SNARKV_OK;
}

int snarkv_g1_msm_plan(snarkv_ctx* ctx, size_t n, uint32_t out[4]) {
    if (!ctx || !out) return SNARKV_ERR_USAGE;
    msm_plan_query(ctx, n, out);
    return SNARKV_OK;
}

int snarkv_profile_enable(snarkv_ctx* ctx, int on) {
    if (!ctx) return SNARKV_ERR_USAGE;
    ctx->profiling = on != 0;
    return SNARKV_OK;
}

int snarkv_profile_read(snarkv_ctx* ctx, snarkv_stage_time* out, int cap) {
    CTX_GUARD(ctx);
    if (!out || cap < 0) return ctx->fail(SNARKV_ERR_USAGE, "snarkv_profile_read: bad argument");
    SNARKV_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    int k = 0;
    for (const StageRecord& r : ctx->stages) {
        if (k >= cap) break;
        float ms = 0.f;
        SNARKV_CUDA_TRY(ctx, cudaEventElapsedTime(&ms, r.start, r.stop));
        out[k].name = r.name;
        out[k].ms = ms;
        out[k].launches = r.launches;
        ++k;
    }
    return k;
}

uint64_t snarkv_launch_count(const snarkv_ctx* ctx) { return ctx ? ctx->launches : 0; }

// ---- MSM --------------------------------------------------------------------------------------------------------------
int snarkv_g1_msm(snarkv_ctx* ctx, const uint8_t* scalars, const uint8_t* points, size_t n, int format, int flags,
                  uint8_t out_affine[64]) {
    CTX_GUARD(ctx);
    if (n == 0) return ctx->fail(SNARKV_ERR_EMPTY, "multi_scalar_multiplication on an empty slice");
    if (!scalars || !points || !out_affine || bad_format(format)) return ctx->fail(SNARKV_ERR_USAGE, "snarkv_g1_msm: bad argument");
    ctx->profile_begin_call();
    return msm_run_host(ctx, scalars, points, n, format, flags, out_affine, nullptr);
}

// ---- resident base set: the preprocessed / vk commitments of a protocol are fixed, only the scalars change per call ---------
struct snarkv_bases {
    int device;
    size_t n;
    void* d_points;   // 2 n x 64 B: Montgomery P_i, then phi(P_i) for the GLV plans
};

int snarkv_g1_bases_upload(snarkv_ctx* ctx, const uint8_t* points, size_t n, int format, int flags, snarkv_bases** out) {
    CTX_GUARD(ctx);
    if (!out) return ctx->fail(SNARKV_ERR_USAGE, "snarkv_g1_bases_upload: bad argument");
    *out = nullptr;
    if (n == 0) return ctx->fail(SNARKV_ERR_EMPTY, "empty base set");
    if (!points || bad_format(format) || n >= (1ull << 30)) return ctx->fail(SNARKV_ERR_USAGE, "snarkv_g1_bases_upload: bad argument");
    snarkv_bases* b = new (std::nothrow) snarkv_bases();
    if (!b) return ctx->fail(SNARKV_ERR_USAGE, "out of host memory");
    b->device = ctx->device;
    b->n = n;
    cudaError_t ce = cudaMalloc(&b->d_points, 2 * n * 64);
    if (ce != cudaSuccess) { delete b; return ctx->fail(SNARKV_ERR_CUDA, "cudaMalloc(resident bases)", ce); }
    uint8_t* d_in = (uint8_t*)ctx->wsget(WS_IO_B, n * 64);
    int* d_status = (int*)ctx->wsget(WS_STATUS, 4);
    int rc = (!d_in || !d_status) ? SNARKV_ERR_CUDA : SNARKV_OK;
    int status = 0;
    cudaStream_t st = ctx->stream;
    if (!rc && (ce = cudaMemcpyAsync(d_in, points, n * 64, cudaMemcpyHostToDevice, st)) != cudaSuccess) rc = ctx->fail(SNARKV_ERR_CUDA, "H2D", ce);
    if (!rc) rc = msm_bases_prepare(ctx, d_in, n, format, (flags & SNARKV_CHECK_INPUTS) ? 1 : 0, b->d_points, d_status);
    if (!rc && (ce = cudaMemcpyAsync(&status, d_status, 4, cudaMemcpyDeviceToHost, st)) != cudaSuccess) rc = ctx->fail(SNARKV_ERR_CUDA, "D2H", ce);
    if ((ce = cudaStreamSynchronize(st)) != cudaSuccess && !rc) rc = ctx->fail(SNARKV_ERR_CUDA, "cudaStreamSynchronize", ce);
    if (!rc && status != 0) rc = ctx->fail(status, "point is not a valid G1Affine");
    if (rc) {
        cudaFree(b->d_points);
        delete b;
        return rc;
    }
    *out = b;
    return SNARKV_OK;
}

void snarkv_g1_bases_free(snarkv_ctx* ctx, snarkv_bases* bases) {
    if (!bases) return;
    if (ctx) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
    }
    cudaFree(bases->d_points);
    delete bases;
}

int snarkv_g1_msm_bases_resident(snarkv_ctx* ctx, const snarkv_bases* bases, const uint8_t* scalars, size_t n, int format, int flags,
                                 uint8_t out_affine[64]) {
    CTX_GUARD(ctx);
    if (n == 0) return ctx->fail(SNARKV_ERR_EMPTY, "multi_scalar_multiplication on an empty slice");
    if (!bases || !scalars || !out_affine || bad_format(format)) return ctx->fail(SNARKV_ERR_USAGE, "snarkv_g1_msm_bases_resident: bad argument");
    if (bases->device != ctx->device) return ctx->fail(SNARKV_ERR_USAGE, "base set lives on another device");
    if (n != bases->n) return ctx->fail(SNARKV_ERR_USAGE, "one scalar per resident base: n must equal the size of the base set");
    ctx->profile_begin_call();
    return msm_run_host(ctx, scalars, nullptr, n, format, flags, out_affine, nullptr, (const uint8_t*)bases->d_points);
}

int snarkv_g1_msm_partial(snarkv_ctx* ctx, const uint8_t* scalars, const uint8_t* points, size_t n, int format, int flags,
                          void* d_out_jacobian) {
    CTX_GUARD(ctx);
    if (n == 0) return ctx->fail(SNARKV_ERR_EMPTY, "multi_scalar_multiplication on an empty slice");
    if (!scalars || !points || !d_out_jacobian || bad_format(format)) return ctx->fail(SNARKV_ERR_USAGE, "snarkv_g1_msm_partial: bad argument");
    ctx->profile_begin_call();
    return msm_run_host(ctx, scalars, points, n, format, flags, nullptr, d_out_jacobian);
}

int snarkv_g1_msm_device(snarkv_ctx* ctx, const void* d_scalars, const void* d_points, size_t n, int format, int flags,
                         void* d_out_affine, void* d_out_jacobian, void* d_status) {
    CTX_GUARD(ctx);
    if (n == 0) return ctx->fail(SNARKV_ERR_EMPTY, "multi_scalar_multiplication on an empty slice");
    if (!d_scalars || !d_points || bad_format(format)) return ctx->fail(SNARKV_ERR_USAGE, "snarkv_g1_msm_device: bad argument");
    if ((flags & SNARKV_CHECK_INPUTS) && !d_status)
        return ctx->fail(SNARKV_ERR_USAGE, "snarkv_g1_msm_device: SNARKV_CHECK_INPUTS needs a d_status word to report into");
    ctx->profile_begin_call();
    return msm_run_device(ctx, d_scalars, d_points, n, format, format, format, flags, d_out_affine, d_out_jacobian, d_status);
}

int snarkv_g1_fold_partials_device(snarkv_ctx* ctx, const void* d_partials, size_t k, int format, void* d_out_affine) {
    CTX_GUARD(ctx);
    if (!d_partials || !d_out_affine || bad_format(format)) return ctx->fail(SNARKV_ERR_USAGE, "fold_partials: bad argument");
    return msm_fold_partials_device(ctx, d_partials, k, format, d_out_affine);
}

int snarkv_g1_msm_batch(snarkv_ctx* ctx, const uint8_t* scalars, const uint8_t* points, const uint64_t* offsets, size_t m,
                        int format, int flags, uint8_t* out_affine) {
    CTX_GUARD(ctx);
    if (!scalars || !points || !offsets || !out_affine || bad_format(format) || m == 0)
        return ctx->fail(SNARKV_ERR_USAGE, "snarkv_g1_msm_batch: bad argument");
    if (offsets[0] != 0) return ctx->fail(SNARKV_ERR_USAGE, "offsets[0] must be 0");
    for (size_t j = 0; j < m; ++j)
        if (offsets[j + 1] <= offsets[j]) return ctx->fail(SNARKV_ERR_EMPTY, "empty or decreasing MSM segment");
    const size_t total = offsets[m];
    ctx->profile_begin_call();
    uint8_t* d_s = (uint8_t*)ctx->wsget(WS_IO_A, total * 32);
    uint8_t* d_p = (uint8_t*)ctx->wsget(WS_IO_B, total * 64);
    uint64_t* d_off = (uint64_t*)ctx->wsget(WS_IO_C, (m + 1) * 8);
    uint8_t* d_out = (uint8_t*)ctx->wsget(WS_IO_D, m * 64);
    int* d_status = (int*)ctx->wsget(WS_STATUS, 4);
    if (!d_s || !d_p || !d_off || !d_out || !d_status) return SNARKV_ERR_CUDA;
    cudaStream_t st = ctx->stream;
    SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(d_s, scalars, total * 32, cudaMemcpyHostToDevice, st));
    SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(d_p, points, total * 64, cudaMemcpyHostToDevice, st));
    SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(d_off, offsets, (m + 1) * 8, cudaMemcpyHostToDevice, st));
    int rc = msm_batch_device(ctx, d_s, d_p, d_off, m, total, format, flags, d_out, d_status);
    if (rc) return rc;
    int status = 0;
    SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(out_affine, d_out, m * 64, cudaMemcpyDeviceToHost, st));
    SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(&status, d_status, 4, cudaMemcpyDeviceToHost, st));
    SNARKV_CUDA_TRY(ctx, cudaStreamSynchronize(st));
    if (status != 0) return ctx->fail(status, status == SNARKV_ERR_BAD_SCALAR ? "scalar is not a canonical Fr" : "point is not a valid G1Affine");
    return SNARKV_OK;
}

// sum_j rho^j * MSM_j as ONE MSM (scalars scaled on the device)
int snarkv_g1_msm_batch_rlc(snarkv_ctx* ctx, const uint8_t* scalars, const uint8_t* points, const uint64_t* offsets, size_t m,
                            const uint8_t rho[32], int format, int flags, uint8_t out_affine[64]) {
    CTX_GUARD(ctx);
    if (!out_affine || !offsets) return ctx->fail(SNARKV_ERR_USAGE, "snarkv_g1_msm_batch_rlc: bad argument");
    if (offsets[0] != 0) return ctx->fail(SNARKV_ERR_USAGE, "offsets[0] must be 0");
    return msm_batch_rlc_host(ctx, scalars, points, offsets, m, rho, format, flags, 0, out_affine, nullptr);
}

// ---- KzgAs::verify -------------------------------------------------------------------------------------------------------
int snarkv_kzg_accumulate(snarkv_ctx* ctx, const uint8_t* lhs, const uint8_t* rhs, size_t n, const uint8_t r[32], int format,
                          uint8_t out_lhs[64], uint8_t out_rhs[64]) {
    CTX_GUARD(ctx);
    if (n == 0) return ctx->fail(SNARKV_ERR_EMPTY, "accumulate over zero accumulators");
    if (!lhs || !rhs || !r || !out_lhs || !out_rhs || bad_format(format)) return ctx->fail(SNARKV_ERR_USAGE, "snarkv_kzg_accumulate: bad argument");
    ctx->profile_begin_call();
    uint8_t* d_l = (uint8_t*)ctx->wsget(WS_IO_A, n * 64);
    uint8_t* d_r = (uint8_t*)ctx->wsget(WS_IO_B, n * 64);
    uint8_t* d_pw = (uint8_t*)ctx->wsget(WS_IO_C, n * 32 + 32);  // [r | powers]
    uint8_t* d_o = (uint8_t*)ctx->wsget(WS_OUT, 256);   // [lhs' 64 | rhs' 64 | status 4]
    if (!d_l || !d_r || !d_pw || !d_o) return SNARKV_ERR_CUDA;
    cudaStream_t st = ctx->stream;
    SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(d_pw, r, 32, cudaMemcpyHostToDevice, st));
    SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(d_l, lhs, n * 64, cudaMemcpyHostToDevice, st));
    SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(d_r, rhs, n * 64, cudaMemcpyHostToDevice, st));
    int rc = fr_powers_device(ctx, d_pw, format, n, d_pw + 32);
    if (rc) return rc;
    // accumulators come out of proofs: every point is validated like `read_ec_point` / `from_xy` does before the reference ever
    // adds it (canonical coordinates, on the curve) — the a = 0 addition formulas would accept an off-curve point silently
    rc = msm_run_device_pair(ctx, d_pw + 32, d_l, d_r, n, SNARKV_MONTGOMERY, format, format, SNARKV_CHECK_INPUTS, d_o, d_o + 128);
    if (rc) return rc;
    uint8_t host[132];
    SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(host, d_o, 132, cudaMemcpyDeviceToHost, st));
    SNARKV_CUDA_TRY(ctx, cudaStreamSynchronize(st));
    int status;
    memcpy(&status, host + 128, 4);
    if (status != 0) return ctx->fail(status, "accumulator point is not a valid G1Affine");
    memcpy(out_lhs, host, 64);
    memcpy(out_rhs, host + 64, 64);
    return SNARKV_OK;
}

// ---- KZG decide --------------------------------------------------------------------------------------------------------
int snarkv_kzg_set_deciding_key(snarkv_ctx* ctx, const uint8_t g1[64], const uint8_t g2[128], const uint8_t s_g2[128]) {
    CTX_GUARD(ctx);
    if (!g1 || !g2 || !s_g2) return ctx->fail(SNARKV_ERR_USAGE, "snarkv_kzg_set_deciding_key: null argument");
    return kzg_set_key(ctx, g1, g2, s_g2);
}

int snarkv_kzg_decide_batch_device(snarkv_ctx* ctx, const void* d_lhs, const void* d_rhs, size_t N, int format, void* d_accept,
                                   void* d_gt) {
    CTX_GUARD(ctx);
    if (!ctx->has_key) return ctx->fail(SNARKV_ERR_NO_KEY, "decide before snarkv_kzg_set_deciding_key");
    if (!d_lhs || !d_rhs || !d_accept || bad_format(format)) return ctx->fail(SNARKV_ERR_USAGE, "snarkv_kzg_decide_batch_device: bad argument");
    if (N == 0) return SNARKV_OK;
    ctx->profile_begin_call();
    return kzg_decide_device(ctx, d_lhs, d_rhs, N, format, d_accept, d_gt);
}

int snarkv_kzg_decide_batch(snarkv_ctx* ctx, const uint8_t* lhs, const uint8_t* rhs, size_t N, int format, uint8_t* accept,
                            uint8_t* gt_out) {
    CTX_GUARD(ctx);
    if (!ctx->has_key) return ctx->fail(SNARKV_ERR_NO_KEY, "decide before snarkv_kzg_set_deciding_key");
    if (!lhs || !rhs || !accept || bad_format(format)) return ctx->fail(SNARKV_ERR_USAGE, "snarkv_kzg_decide_batch: bad argument");
    if (N == 0) return SNARKV_OK;  // decide_all over an empty Vec is Ok(())
    ctx->profile_begin_call();
    uint8_t* d_l = (uint8_t*)ctx->wsget(WS_IO_A, N * 64);
    uint8_t* d_r = (uint8_t*)ctx->wsget(WS_IO_B, N * 64);
    uint8_t* d_acc = (uint8_t*)ctx->wsget(WS_IO_C, N);
    uint8_t* d_gt = gt_out ? (uint8_t*)ctx->wsget(WS_IO_D, N * 384) : nullptr;
    if (!d_l || !d_r || !d_acc || (gt_out && !d_gt)) return SNARKV_ERR_CUDA;
    cudaStream_t st = ctx->stream;
    SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(d_l, lhs, N * 64, cudaMemcpyHostToDevice, st));
    SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(d_r, rhs, N * 64, cudaMemcpyHostToDevice, st));
    int rc = kzg_decide_device(ctx, d_l, d_r, N, format, d_acc, d_gt);
    if (rc) return rc;
    SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(accept, d_acc, N, cudaMemcpyDeviceToHost, st));
    if (gt_out) SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(gt_out, d_gt, N * 384, cudaMemcpyDeviceToHost, st));
    SNARKV_CUDA_TRY(ctx, cudaStreamSynchronize(st));
    return SNARKV_OK;
}

// ---- fused batch decision (RLC) ---------------------------------------------------------------------------------------------
int snarkv_kzg_decide_all_fused(snarkv_ctx* ctx, const uint8_t* lhs, const uint8_t* rhs, size_t N, const uint8_t rho[32], int format,
                                uint8_t* accept, uint8_t out_lhs[64], uint8_t out_rhs[64]) {
    CTX_GUARD(ctx);
    if (!ctx->has_key) return ctx->fail(SNARKV_ERR_NO_KEY, "decide before snarkv_kzg_set_deciding_key");
    if (!lhs || !rhs || !rho || !accept || bad_format(format)) return ctx->fail(SNARKV_ERR_USAGE, "snarkv_kzg_decide_all_fused: bad argument");
    if (N == 0) { *accept = 1; return SNARKV_OK; }   // decide_all over an empty Vec is Ok(())
    ctx->profile_begin_call();
    uint8_t* d_l = (uint8_t*)ctx->wsget(WS_IO_A, N * 64);
    uint8_t* d_r = (uint8_t*)ctx->wsget(WS_IO_B, N * 64);
    uint8_t* d_pw = (uint8_t*)ctx->wsget(WS_IO_C, N * 32 + 32);
    uint8_t* d_o = (uint8_t*)ctx->wsget(WS_OUT, 256);   // [lhs' | rhs' | accept]
    if (!d_l || !d_r || !d_pw || !d_o) return SNARKV_ERR_CUDA;
    cudaStream_t st = ctx->stream;
    SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(d_pw, rho, 32, cudaMemcpyHostToDevice, st));
    SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(d_l, lhs, N * 64, cudaMemcpyHostToDevice, st));
    SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(d_r, rhs, N * 64, cudaMemcpyHostToDevice, st));
    int rc = fr_powers_device(ctx, d_pw, format, N, d_pw + 32);
    if (rc) return rc;
    // layout of d_o: [lhs' 64 | rhs' 64 | accept 1 .. pad to 132 | status 4]  (one status word covers both base sets)
    rc = msm_run_device_pair(ctx, d_pw + 32, d_l, d_r, N, SNARKV_MONTGOMERY, format, format, SNARKV_CHECK_INPUTS, d_o, d_o + 132);
    if (rc) return rc;
    rc = kzg_decide_device(ctx, d_o, d_o + 64, 1, format, d_o + 128, nullptr);
    if (rc) return rc;
    uint8_t host[136];
    SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(host, d_o, 136, cudaMemcpyDeviceToHost, st));
    SNARKV_CUDA_TRY(ctx, cudaStreamSynchronize(st));
    int status;
    memcpy(&status, host + 132, 4);
    *accept = (status == 0 && host[128] == 1) ? 1 : 0;   // a non-curve accumulator rejects the batch
    if (out_lhs) memcpy(out_lhs, host, 64);
    if (out_rhs) memcpy(out_rhs, host + 64, 64);
    return SNARKV_OK;
}

// ---- (f1) Fr scalar preparation ------------------------------------------------------------------------------------------------
int snarkv_fr_powers(snarkv_ctx* ctx, const uint8_t r[32], size_t n, int format, uint8_t* out) {
    CTX_GUARD(ctx);
    if (!r || !out || bad_format(format)) return ctx->fail(SNARKV_ERR_USAGE, "snarkv_fr_powers: bad argument");
    if (n == 0) return SNARKV_OK;
    uint8_t* d = (uint8_t*)ctx->wsget(WS_IO_A, n * 32 + 32);
    if (!d) return SNARKV_ERR_CUDA;
    cudaStream_t st = ctx->stream;
    SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(d, r, 32, cudaMemcpyHostToDevice, st));
    int rc = fr_powers_device(ctx, d, format, n, d + 32);
    if (rc) return rc;
    if (format == SNARKV_CANONICAL) {
        rc = fr_from_mont_device(ctx, d + 32, n);
        if (rc) return rc;
    }
    SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(out, d + 32, n * 32, cudaMemcpyDeviceToHost, st));
    SNARKV_CUDA_TRY(ctx, cudaStreamSynchronize(st));
    return SNARKV_OK;
}

int snarkv_fr_batch_invert(snarkv_ctx* ctx, uint8_t* values, size_t n, const uint8_t* coeff, int format) {
    CTX_GUARD(ctx);
    if (!values || bad_format(format)) return ctx->fail(SNARKV_ERR_USAGE, "snarkv_fr_batch_invert: bad argument");
    if (n == 0) return SNARKV_OK;
    uint8_t* d_v = (uint8_t*)ctx->wsget(WS_IO_A, n * 32);
    uint8_t* d_s = (uint8_t*)ctx->wsget(WS_IO_B, n * 32);
    uint8_t* d_c = (uint8_t*)ctx->wsget(WS_IO_C, 32);
    if (!d_v || !d_s || !d_c) return SNARKV_ERR_CUDA;
    cudaStream_t st = ctx->stream;
    SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(d_v, values, n * 32, cudaMemcpyHostToDevice, st));
    if (coeff) SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(d_c, coeff, 32, cudaMemcpyHostToDevice, st));
    int rc = fr_batch_invert_device(ctx, d_v, n, format, coeff ? d_c : nullptr, d_s);
    if (rc) return rc;
    SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(values, d_v, n * 32, cudaMemcpyDeviceToHost, st));
    SNARKV_CUDA_TRY(ctx, cudaStreamSynchronize(st));
    return SNARKV_OK;
}

int snarkv_fr_mul_vec(snarkv_ctx* ctx, const uint8_t* a, const uint8_t* b, size_t n, int format, uint8_t* out) {
    CTX_GUARD(ctx);
    if (!a || !b || !out || bad_format(format)) return ctx->fail(SNARKV_ERR_USAGE, "snarkv_fr_mul_vec: bad argument");
    if (n == 0) return SNARKV_OK;
    uint8_t* d_a = (uint8_t*)ctx->wsget(WS_IO_A, n * 32);
    uint8_t* d_b = (uint8_t*)ctx->wsget(WS_IO_B, n * 32);
    uint8_t* d_o = (uint8_t*)ctx->wsget(WS_IO_C, n * 32);
    if (!d_a || !d_b || !d_o) return SNARKV_ERR_CUDA;
    cudaStream_t st = ctx->stream;
    SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(d_a, a, n * 32, cudaMemcpyHostToDevice, st));
    SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(d_b, b, n * 32, cudaMemcpyHostToDevice, st));
    int rc = fr_mul_vec_device(ctx, d_a, d_b, n, format, d_o);
    if (rc) return rc;
    SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(out, d_o, n * 32, cudaMemcpyDeviceToHost, st));
    SNARKV_CUDA_TRY(ctx, cudaStreamSynchronize(st));
    return SNARKV_OK;
}

// ---- (a13) LimbsEncoding::from_repr for a batch of accumulators (pcs/kzg/accumulator.rs:57-81) ------------------------------------
int snarkv_kzg_accumulators_from_limbs(snarkv_ctx* ctx, const uint8_t* limbs, size_t m, uint32_t num_limbs, uint32_t limb_bits, int format,
                                       uint8_t* lhs, uint8_t* rhs, uint8_t* valid) {
    CTX_GUARD(ctx);
    if (m == 0 && !bad_format(format)) return SNARKV_OK;   // an empty batch needs no buffers
    if (!limbs || !lhs || !rhs || !valid || bad_format(format) || num_limbs == 0 || num_limbs > 8 || limb_bits == 0 ||
        (uint64_t)limb_bits * (num_limbs - 1) > 256)
        return ctx->fail(SNARKV_ERR_USAGE, "snarkv_kzg_accumulators_from_limbs: bad argument");
    if (m == 0) return SNARKV_OK;
    const size_t in_bytes = m * 4 * num_limbs * 32;
    uint8_t* d_in = (uint8_t*)ctx->wsget(WS_IO_A, in_bytes);
    uint8_t* d_l = (uint8_t*)ctx->wsget(WS_IO_B, m * 64);
    uint8_t* d_r = (uint8_t*)ctx->wsget(WS_IO_C, m * 64);
    uint8_t* d_v = (uint8_t*)ctx->wsget(WS_IO_D, m);
    if (!d_in || !d_l || !d_r || !d_v) return SNARKV_ERR_CUDA;
    cudaStream_t st = ctx->stream;
    SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(d_in, limbs, in_bytes, cudaMemcpyHostToDevice, st));
    int rc = accumulators_from_limbs_device(ctx, d_in, m, num_limbs, limb_bits, format, d_l, d_r, d_v);
    if (rc) return rc;
    SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(lhs, d_l, m * 64, cudaMemcpyDeviceToHost, st));
    SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(rhs, d_r, m * 64, cudaMemcpyDeviceToHost, st));
    SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(valid, d_v, m, cudaMemcpyDeviceToHost, st));
    SNARKV_CUDA_TRY(ctx, cudaStreamSynchronize(st));
    return SNARKV_OK;
}

// ---- (f3) straight-line Fr program for a batch of proofs (protocol.rs:211-283, 333-392; proof.rs:298-349) -------------------------
// The program is small and comes from the host compiler: validate it here (opcodes, register bounds, write-before-read), then stage
// instructions | consts | out_regs in one device buffer.
static int fr_program_stage(snarkv_ctx* ctx, const snarkv_fr_instr* program, size_t n_instr, uint32_t n_regs, const uint8_t* consts,
                            size_t n_consts, size_t n_inputs, size_t m, const uint32_t* out_regs, size_t n_out, int format,
                            uint8_t** d_prog, uint8_t** d_consts, uint8_t** d_out_regs) {
    if (!program || n_instr == 0 || n_instr > (1u << 24) || n_regs == 0 || n_regs > (1u << 20) || !out_regs || n_out == 0 || bad_format(format) ||
        (n_consts && !consts) || n_consts > (1u << 24) || n_inputs > (1u << 24) || m > ((size_t)1 << 31))
        return ctx->fail(SNARKV_ERR_USAGE, "snarkv_fr_program_eval_batch: bad argument");
    std::vector<uint8_t> written(n_regs, 0);
    for (size_t i = 0; i < n_instr; ++i) {
        const snarkv_fr_instr& in = program[i];
        bool ok = in.dst < n_regs;
        auto src = [&](uint32_t r) { return r < n_regs && written[r]; };
        switch (in.op) {
            case SNARKV_FR_OP_INPUT: ok = ok && in.a < n_inputs; break;
            case SNARKV_FR_OP_CONST: ok = ok && in.a < n_consts; break;
            case SNARKV_FR_OP_ADD: case SNARKV_FR_OP_SUB: case SNARKV_FR_OP_MUL: case SNARKV_FR_OP_KEEPZ: ok = ok && src(in.a) && src(in.b); break;
            case SNARKV_FR_OP_NEG: case SNARKV_FR_OP_INV: case SNARKV_FR_OP_NZ: ok = ok && src(in.a); break;
            default: ok = false;
        }
        if (!ok) return ctx->fail(SNARKV_ERR_USAGE, "snarkv_fr_program_eval_batch: invalid instruction (opcode, operand out of range, or register read before it is written)");
        written[in.dst] = 1;
    }
    for (size_t k = 0; k < n_out; ++k)
        if (out_regs[k] >= n_regs || !written[out_regs[k]]) return ctx->fail(SNARKV_ERR_USAGE, "snarkv_fr_program_eval_batch: output register never written");
    const size_t prog_bytes = n_instr * sizeof(snarkv_fr_instr), const_bytes = n_consts * 32, out_bytes = n_out * 4;
    uint8_t* buf = (uint8_t*)ctx->wsget(WS_FR_PROG, prog_bytes + const_bytes + out_bytes + 64);
    if (!buf) return SNARKV_ERR_CUDA;
    cudaStream_t st = ctx->stream;
    *d_prog = buf;
    *d_consts = buf + prog_bytes;          // 16-byte aligned: prog_bytes is a multiple of 16
    *d_out_regs = *d_consts + const_bytes;
    SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(*d_prog, program, prog_bytes, cudaMemcpyHostToDevice, st));
    if (n_consts) SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(*d_consts, consts, const_bytes, cudaMemcpyHostToDevice, st));
    SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(*d_out_regs, out_regs, out_bytes, cudaMemcpyHostToDevice, st));
    return SNARKV_OK;
}

int snarkv_fr_program_eval_batch_device(snarkv_ctx* ctx, const snarkv_fr_instr* program, size_t n_instr, uint32_t n_regs,
                                        const uint8_t* consts, size_t n_consts, const void* d_inputs, size_t n_inputs, size_t m,
                                        const uint32_t* out_regs, size_t n_out, int format, void* d_outputs) {
    CTX_GUARD(ctx);
    if (m != 0 && (!d_outputs || (n_inputs && !d_inputs))) return ctx->fail(SNARKV_ERR_USAGE, "snarkv_fr_program_eval_batch_device: bad argument");
    uint8_t *d_prog, *d_consts, *d_out_regs;
    int rc = fr_program_stage(ctx, program, n_instr, n_regs, consts, n_consts, n_inputs, m, out_regs, n_out, format, &d_prog, &d_consts, &d_out_regs);
    if (rc) return rc;
    if (m == 0) return SNARKV_OK;
    rc = fr_program_device(ctx, d_prog, n_instr, d_consts, n_consts, d_inputs, n_inputs, m, format, n_regs, d_out_regs, n_out, d_outputs);
    if (rc) return rc;
    // the staged program is pageable host memory handed to cudaMemcpyAsync: make sure it has left the caller's buffers
    SNARKV_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return SNARKV_OK;
}

int snarkv_fr_program_eval_batch(snarkv_ctx* ctx, const snarkv_fr_instr* program, size_t n_instr, uint32_t n_regs, const uint8_t* consts,
                                 size_t n_consts, const uint8_t* inputs, size_t n_inputs, size_t m, const uint32_t* out_regs, size_t n_out,
                                 int format, uint8_t* outputs) {
    CTX_GUARD(ctx);
    if (m != 0 && (!outputs || (n_inputs && !inputs))) return ctx->fail(SNARKV_ERR_USAGE, "snarkv_fr_program_eval_batch: bad argument");
    uint8_t *d_prog, *d_consts, *d_out_regs;
    int rc = fr_program_stage(ctx, program, n_instr, n_regs, consts, n_consts, n_inputs, m, out_regs, n_out, format, &d_prog, &d_consts, &d_out_regs);
    if (rc) return rc;
    if (m == 0) return SNARKV_OK;
    uint8_t* d_in = (uint8_t*)ctx->wsget(WS_IO_A, m * n_inputs * 32 + 32);
    uint8_t* d_o = (uint8_t*)ctx->wsget(WS_IO_B, m * n_out * 32);
    if (!d_in || !d_o) return SNARKV_ERR_CUDA;
    cudaStream_t st = ctx->stream;
    if (n_inputs) SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(d_in, inputs, m * n_inputs * 32, cudaMemcpyHostToDevice, st));
    rc = fr_program_device(ctx, d_prog, n_instr, d_consts, n_consts, d_in, n_inputs, m, format, n_regs, d_out_regs, n_out, d_o);
    if (rc) return rc;
    SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(outputs, d_o, m * n_out * 32, cudaMemcpyDeviceToHost, st));
    SNARKV_CUDA_TRY(ctx, cudaStreamSynchronize(st));
    return SNARKV_OK;
}

// ---- (f2) EvmTranscript challenges for a batch of proofs ----------------------------------------------------------------------
int snarkv_evm_transcript_challenges(snarkv_ctx* ctx, const uint8_t* streams, size_t stream_len, const uint32_t* seg_end, size_t k, size_t m,
                                     int format, uint8_t* challenges) {
    CTX_GUARD(ctx);
    if (!seg_end || !challenges || bad_format(format) || k == 0 || (stream_len && !streams) || (stream_len & 31))
        return ctx->fail(SNARKV_ERR_USAGE, "snarkv_evm_transcript_challenges: bad argument");
    uint32_t prev = 0;
    for (size_t i = 0; i < k; ++i) {
        if (seg_end[i] < prev || (seg_end[i] & 31) || seg_end[i] > stream_len) return ctx->fail(SNARKV_ERR_USAGE, "seg_end must be non-decreasing multiples of 32 within the stream");
        prev = seg_end[i];
    }
    if (m == 0) return SNARKV_OK;
    uint8_t* d_s = (uint8_t*)ctx->wsget(WS_IO_A, m * stream_len);
    uint32_t* d_e = (uint32_t*)ctx->wsget(WS_IO_B, k * 4);
    uint8_t* d_o = (uint8_t*)ctx->wsget(WS_IO_C, m * k * 32);
    if (!d_s || !d_e || !d_o) return SNARKV_ERR_CUDA;
    cudaStream_t st = ctx->stream;
    if (stream_len) SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(d_s, streams, m * stream_len, cudaMemcpyHostToDevice, st));
    SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(d_e, seg_end, k * 4, cudaMemcpyHostToDevice, st));
    int rc = evm_transcript_device(ctx, d_s, stream_len, d_e, k, m, format, d_o);
    if (rc) return rc;
    SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(challenges, d_o, m * k * 32, cudaMemcpyDeviceToHost, st));
    SNARKV_CUDA_TRY(ctx, cudaStreamSynchronize(st));
    return SNARKV_OK;
}

// ---- test support ---------------------------------------------------------------------------------------------------------
int snarkv_debug_field_op(snarkv_ctx* ctx, int field, int op, const uint8_t* a, const uint8_t* b, size_t n, uint8_t* out) {
    CTX_GUARD(ctx);
    if (!a || !b || !out || n == 0 || field < 0 || field > 1) return ctx->fail(SNARKV_ERR_USAGE, "snarkv_debug_field_op: bad argument");
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

// ---- synthetic workload ---------------------------------------------------------------------------------------------------
int snarkv_synth_scalars_device(snarkv_ctx* ctx, uint64_t seed, uint64_t start, size_t n, int format, void* d_out) {
    CTX_GUARD(ctx);
    if (!d_out || bad_format(format)) return ctx->fail(SNARKV_ERR_USAGE, "snarkv_synth_scalars_device: bad argument");
    return synth_scalars_device(ctx, seed, start, n, format, d_out);
}
int snarkv_synth_points_device(snarkv_ctx* ctx, uint64_t seed, uint64_t start, size_t n, int format, void* d_out) {
    CTX_GUARD(ctx);
    if (!d_out || bad_format(format)) return ctx->fail(SNARKV_ERR_USAGE, "snarkv_synth_points_device: bad argument");
    return synth_points_device(ctx, seed, start, n, format, d_out);
}

}  // extern "C"
