// multi.cu — one host call, all the GPUs of the box: the multi-device entry points of include/snarkv_cuda.h.
//
// The reference parallelises its large MSM by cutting the term slice into one contiguous chunk per rayon thread and folding the
// per-chunk results (snark-verifier/src/util/msm.rs:322-336).  `snarkv_multi_g1_msm` is that shape with GPUs for threads: device g
// gets terms [g * ceil(n / G), (g + 1) * ceil(n / G)), runs the whole single-device pipeline on them (msm.cu) from its own host
// thread and leaves a 96-byte Jacobian partial in its HBM; the ONE exchange step — the fold of util/msm.rs:333-335 — is a single
// kernel on device 0 that reads every partial straight out of its producer's memory over NVLink peer access (k_fold_partials_peer:
// gather fused into the fold, no staging copy, no separate collective).  Independent units shard with no exchange at all:
// pairing checks (pcs/kzg/decider.rs:84-93) go to devices in contiguous blocks, RLC-fused batches of small MSMs by segment with
// device g starting its powers of rho at rho^(first segment of g).
//
// A process that prefers one rank per GPU (bench.py under torchrun) uses the per-device entry points + an NCCL all-gather of the
// same 96-byte partials instead; both paths give identical bytes.
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "ctx.hpp"

using namespace snarkv;

struct snarkv_multi {
    int n = 0;
    snarkv_ctx* ctx[SNARKV_MAX_DEVICES] = {};
    void* d_partial[SNARKV_MAX_DEVICES] = {};   // 96 B Jacobian partial in each device's HBM
    void* d_out = nullptr;                      // 64 B on device 0
    std::string err;
    int fail(int code, const std::string& what) {
        err = what;
        return code;
    }
};

namespace {
// run fn(g) for every g < G on its own host thread (each thread binds its device); returns the first non-zero code
template <typename F>
int fan_out(snarkv_multi* m, int G, F fn) {
    std::vector<int> rc(G, 0);
    std::vector<std::thread> pool;
    pool.reserve(G);
    for (int g = 1; g < G; ++g)
        pool.emplace_back([&, g] {
            cudaSetDevice(m->ctx[g]->device);
            rc[g] = fn(g);
        });
    cudaSetDevice(m->ctx[0]->device);
    rc[0] = fn(0);
    for (auto& t : pool) t.join();
    for (int g = 0; g < G; ++g)
        if (rc[g] != 0) {
            m->err = "device " + std::to_string(m->ctx[g]->device) + ": " + m->ctx[g]->err;
            return rc[g];
        }
    return 0;
}
}  // namespace

extern "C" {

int snarkv_multi_init(const int* devices, int n_devices, snarkv_multi** out) {
    if (!out) return SNARKV_ERR_USAGE;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return SNARKV_ERR_CUDA;   // no GPU: fail loudly, no fallback
    if (n_devices <= 0) n_devices = count;                                                  // "all the GPUs of the box"
    if (n_devices > SNARKV_MAX_DEVICES || (!devices && n_devices > count)) return SNARKV_ERR_USAGE;
    snarkv_multi* m = new (std::nothrow) snarkv_multi();
    if (!m) return SNARKV_ERR_USAGE;
    for (int g = 0; g < n_devices; ++g) {
        const int dev = devices ? devices[g] : g;
        int rc = snarkv_init(dev, &m->ctx[g]);
        if (rc == SNARKV_OK && cudaMalloc(&m->d_partial[g], 96) != cudaSuccess) rc = SNARKV_ERR_CUDA;
        if (rc != SNARKV_OK) {
            m->n = g + 1;
            snarkv_multi_destroy(m);
            return rc;
        }
        m->n = g + 1;
    }
    // device 0 reads the other devices' partials directly (NVLink peer access)
    cudaSetDevice(m->ctx[0]->device);
    for (int g = 1; g < m->n; ++g) {
        if (m->ctx[g]->device == m->ctx[0]->device) continue;   // the same GPU listed twice (tests on a 1-GPU box): already local
        int can = 0;
        cudaDeviceCanAccessPeer(&can, m->ctx[0]->device, m->ctx[g]->device);
        cudaError_t ce = can ? cudaDeviceEnablePeerAccess(m->ctx[g]->device, 0) : cudaErrorPeerAccessUnsupported;
        if (ce == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); ce = cudaSuccess; }
        if (ce != cudaSuccess) {
            snarkv_multi_destroy(m);
            return SNARKV_ERR_CUDA;
        }
    }
    if (cudaMalloc(&m->d_out, 64) != cudaSuccess) {
        snarkv_multi_destroy(m);
        return SNARKV_ERR_CUDA;
    }
    *out = m;
    return SNARKV_OK;
}

void snarkv_multi_destroy(snarkv_multi* m) {
    if (!m) return;
    for (int g = 0; g < m->n; ++g) {
        if (!m->ctx[g]) continue;
        cudaSetDevice(m->ctx[g]->device);
        if (m->d_partial[g]) cudaFree(m->d_partial[g]);
        if (g == 0 && m->d_out) cudaFree(m->d_out);
        snarkv_destroy(m->ctx[g]);
    }
    delete m;
}

int snarkv_multi_device_count(const snarkv_multi* m) { return m ? m->n : 0; }
snarkv_ctx* snarkv_multi_ctx(snarkv_multi* m, int i) { return (m && i >= 0 && i < m->n) ? m->ctx[i] : nullptr; }
const char* snarkv_multi_last_error(const snarkv_multi* m) { return m ? m->err.c_str() : "null context"; }

// fold the partials of devices [0, G) on device 0 and copy the affine result out
static int multi_fold(snarkv_multi* m, int G, int format, uint8_t out_affine[64]) {
    snarkv_ctx* c0 = m->ctx[0];
    cudaSetDevice(c0->device);
    int rc = msm_fold_partials_peer(c0, (const void* const*)m->d_partial, (size_t)G, format, m->d_out);
    if (rc) return m->fail(rc, c0->err);
    cudaError_t ce = cudaMemcpyAsync(out_affine, m->d_out, 64, cudaMemcpyDeviceToHost, c0->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(c0->stream);
    if (ce != cudaSuccess) return m->fail(c0->fail(SNARKV_ERR_CUDA, "multi fold", ce), c0->err);
    return SNARKV_OK;
}

int snarkv_multi_g1_msm(snarkv_multi* m, const uint8_t* scalars, const uint8_t* points, size_t n, int format, int flags,
                        uint8_t out_affine[64]) {
    if (!m) return SNARKV_ERR_USAGE;
    m->err.clear();
    if (n == 0) return m->fail(SNARKV_ERR_EMPTY, "multi_scalar_multiplication on an empty slice");
    if (!scalars || !points || !out_affine || (format != SNARKV_CANONICAL && format != SNARKV_MONTGOMERY))
        return m->fail(SNARKV_ERR_USAGE, "snarkv_multi_g1_msm: bad argument");
    // util/msm.rs:322: chunk_size = ceil(n / threads); short inputs use fewer devices (msm.rs:313: fewer terms than threads -> serial)
    const size_t chunk = (n + m->n - 1) / m->n;
    const int G = (int)((n + chunk - 1) / chunk);
    int rc = fan_out(m, G, [&](int g) {
        const size_t lo = (size_t)g * chunk, len = (lo + chunk <= n) ? chunk : n - lo;
        snarkv_ctx* c = m->ctx[g];
        c->err.clear();
        c->profile_begin_call();
        return msm_run_host(c, scalars + lo * 32, points + lo * 64, len, format, flags, nullptr, m->d_partial[g]);
    });
    if (rc) return rc;
    return multi_fold(m, G, format, out_affine);
}

int snarkv_multi_g1_msm_batch_rlc(snarkv_multi* m, const uint8_t* scalars, const uint8_t* points, const uint64_t* offsets, size_t segs,
                                  const uint8_t rho[32], int format, int flags, uint8_t out_affine[64]) {
    if (!m) return SNARKV_ERR_USAGE;
    m->err.clear();
    if (!scalars || !points || !offsets || !rho || !out_affine || segs == 0 || (format != SNARKV_CANONICAL && format != SNARKV_MONTGOMERY))
        return m->fail(SNARKV_ERR_USAGE, "snarkv_multi_g1_msm_batch_rlc: bad argument");
    if (offsets[0] != 0) return m->fail(SNARKV_ERR_USAGE, "offsets[0] must be 0");
    const size_t chunk = (segs + m->n - 1) / m->n;
    const int G = (int)((segs + chunk - 1) / chunk);
    int rc = fan_out(m, G, [&](int g) {
        const size_t j0 = (size_t)g * chunk, cnt = (j0 + chunk <= segs) ? chunk : segs - j0;
        snarkv_ctx* c = m->ctx[g];
        c->err.clear();
        // segment j of the whole batch is scaled by rho^j: this device's powers start at rho^j0
        return msm_batch_rlc_host(c, scalars, points, offsets + j0, cnt, rho, format, flags, (uint64_t)j0, nullptr, m->d_partial[g]);
    });
    if (rc) return rc;
    return multi_fold(m, G, format, out_affine);
}

int snarkv_multi_kzg_set_deciding_key(snarkv_multi* m, const uint8_t g1[64], const uint8_t g2[128], const uint8_t s_g2[128]) {
    if (!m) return SNARKV_ERR_USAGE;
    m->err.clear();
    for (int g = 0; g < m->n; ++g) {
        int rc = snarkv_kzg_set_deciding_key(m->ctx[g], g1, g2, s_g2);
        if (rc) return m->fail(rc, m->ctx[g]->err);
    }
    return SNARKV_OK;
}

int snarkv_multi_kzg_decide_batch(snarkv_multi* m, const uint8_t* lhs, const uint8_t* rhs, size_t N, int format, uint8_t* accept,
                                  uint8_t* gt_out) {
    if (!m) return SNARKV_ERR_USAGE;
    m->err.clear();
    if (N == 0) return SNARKV_OK;
    if (!lhs || !rhs || !accept) return m->fail(SNARKV_ERR_USAGE, "snarkv_multi_kzg_decide_batch: bad argument");
    const size_t chunk = (N + m->n - 1) / m->n;
    const int G = (int)((N + chunk - 1) / chunk);
    return fan_out(m, G, [&](int g) {
        const size_t lo = (size_t)g * chunk, len = (lo + chunk <= N) ? chunk : N - lo;
        return snarkv_kzg_decide_batch(m->ctx[g], lhs + lo * 64, rhs + lo * 64, len, format, accept + lo, gt_out ? gt_out + lo * 384 : nullptr);
    });
}

}  // extern "C"
