// msm.cu — BN254 G1 multi-scalar multiplication on sm_100a: a signed-digit Pippenger pipeline.
//
// Replaces, behind `snarkv_g1_msm*` (include/snarkv_cuda.h), the reference's
//   NativeLoader::multi_scalar_multiplication          snark-verifier/src/loader/native.rs:61-71   (semantics: affine sum)
//   util::msm::multi_scalar_multiplication{,_serial}   snark-verifier/src/util/msm.rs:259-343      (bucket method)
// The reference's bucket method uses unsigned c-bit windows with c = ceil(ln n) + 2 and 2^c - 1 buckets per window,
// processed window by window on one core (or one chunk per rayon thread).  This is a different program with the same
// output: all windows at once, signed digits (2^(c-1) buckets per window), and a counting sort so that each bucket's
// points are contiguous:
//
//   K1 digits_count       scalar -> W signed c-bit digits (stored window-major) + histogram     (HBM + L2 atomics)
//   K2 scan               per-window exclusive prefix sum of the histogram                      (tiny)
//   K3 digits_scatter     window-major: sorted[w][pos] = term index | sign, per-window L2-resident (L2-atomic bound)
//   K4 bucket_accumulate  one thread per (window, bucket): XYZZ += gathered affine points       (IMAD-pipe bound: the MSM)
//   K5 bucket_reduce      per window  sum_b b * B_b  by segment running sums                    (small)
//   K6 window_sum / final Horner over windows with c doublings each, to_affine                  (latency, 1 thread)
//
// HBM layout: scalars n x 32 B and affine points n x 64 B exactly as the caller's slices (halo2curves layout), plus
// workspace: histogram/offset/cursor W x NB x 4 B each, sorted indices W x n x 4 B, bucket sums W x NB x 128 B (XYZZ).
#include "ctx.hpp"
#include "g1.cuh"
#include "glv.cuh"
#include "bucket_affine.cuh"
#include "bucket_chain.cuh"
#include "sort.cuh"

namespace snarkv {

#define SNARKV_HOST_CHUNKS_MAX 7

struct MsmPlan {
    uint32_t c;    // window bits
    uint32_t W;    // windows = ceil(255 / c): the top window absorbs the last signed-digit carry (scalars < 2^254)
    uint32_t NB;   // buckets per window = 2^(c-1), bucket values 1..NB
    uint32_t seg;  // buckets per reduce thread
    uint32_t J;    // segments per window
    uint32_t T;    // max points per accumulate task (a bucket with more points is split into ceil(cnt / T) tasks)
    uint32_t cap;  // task slots per window: sum_b ceil(cnt_b / T) <= NB + n / T
    uint32_t glv;  // 1: every term is split into two half-length terms (P, k1), (phi(P), k2) — glv.cuh
    uint32_t sort_P;        // two-level sort (sort.cuh): coarse partitions per window (0 = small-input path: histogram + scatter)
    uint32_t sort_lo_bits;  // low digit bits resolved inside a partition: NB = sort_P << sort_lo_bits
};
// With GLV the pipeline sees nv = 2 n "virtual terms" (index i < n: P_i with k1, index n + i: phi(P_i) with k2) whose scalars
// have 128 bits, so W = ceil(130 / c) windows instead of ceil(255 / c): the additions are the same in number, but the Horner
// chain, the bucket reduction and the window sums are halved.  Worth it while those fixed costs matter (n < 2^22); above, a
// larger window (c = 17, 15 windows) with long enough bucket lists for the batched-affine kernel saves more than GLV saves tail.
static inline size_t plan_virtual_terms(const MsmPlan& p, size_t n) { return p.glv ? 2 * n : n; }

static int choose_window_bits(size_t n, bool glv) {
    // Measured sweeps on B200 (profiles/r01_window_sweep_*.txt, r01_window_sweep_glv.txt).  n = number of (virtual) terms.
    // Window sizes whose TOP window holds only a few significant bits funnel a large share of the terms into a handful of
    // counters/buckets and are avoided: for 254-bit scalars that rules out c = 9, 11, 12, 14, 18; for the 128-bit halves of
    // the GLV split the good sizes are the ones dividing 128 (8, 16) or leaving an almost empty top window (13).
    if (glv) {
        if (n < (1u << 12)) return 8;
        if (n < (1u << 17)) return 13;
        return 16;
    }
    if (n < (1u << 8)) return 6;
    if (n < (1u << 12)) return 8;
    if (n < (1u << 17)) return 13;
    if (n < (1u << 20)) return 15;
    if (n < (1u << 22)) return 16;
    return 17;
}

static MsmPlan make_plan(size_t n_terms, int c_override, int glv_mode) {
    MsmPlan p;
    // measured plan table (tools/plan_sweep.py, profiles/r02_plan_sweep.txt): GLV wins up to 2^21 terms; from 2^22 on the plain
    // c = 17 plan with the batched-affine kernel is faster (2^22: 12.0 vs 12.5 ms)
    p.glv = (glv_mode == 1 || (glv_mode == 0 && n_terms < ((size_t)1 << 22))) ? 1u : 0u;
#ifdef SNARKV_CURVE_PALLAS
    p.glv = 0;   // glv.cuh holds BN254's lattice; the Pallas build runs the plain pipeline
#endif
    const size_t n = p.glv ? 2 * n_terms : n_terms;
    int c = c_override > 0 ? c_override : choose_window_bits(n, p.glv != 0);
    if (c < 2) c = 2;
    if (c > 22) c = 22;
    p.c = (uint32_t)c;
    p.W = ((p.glv ? 130u : (uint32_t)SNARKV_FIELD_BITS + 1u) + p.c - 1) / p.c;   // scalar bits + the signed-digit carry
    p.NB = 1u << (p.c - 1);
    // bucket-reduce segment length: each thread's chain is 2 seg additions + one small scalar multiple; short segments cut that
    // latency, long ones cut the total work (the small multiples) once there are enough buckets to fill the machine
    const size_t all_buckets = (size_t)p.W * p.NB;
    // (measured with the block-level fold in k_bucket_reduce: 2^18 buckets 16 -> 0.52 ms, 8 -> 0.50, 32 -> 0.68; 2^20 buckets
    // 16 -> 1.26 ms, 32 -> 0.89)
    p.seg = all_buckets >= ((size_t)1 << 19) + ((size_t)1 << 18) ? 32 : all_buckets >= ((size_t)1 << 18) ? 16 : 4;
    if (p.seg > p.NB) p.seg = p.NB;
    p.J = (p.NB + p.seg - 1) / p.seg;
    const size_t avg = (n + p.NB - 1) / p.NB;
    size_t T = 2 * avg;   // uniformly random digits never reach 2x the mean once the mean is >= 64
    if (T < 64) T = 64;
    p.T = (uint32_t)T;
    p.cap = p.NB + (uint32_t)(n / T) + 1;
    // two-level sort for large inputs: partitions of ~32 K references (what one k_sort_buckets block stages in shared memory)
    p.sort_P = 0;
    p.sort_lo_bits = 0;
    if (n >= ((size_t)1 << 18) && p.c >= 9) {
        uint32_t P = 1;
        while ((size_t)P * 32768 < n && P < p.NB) P <<= 1;
        uint32_t lo = (p.c - 1);
        for (uint32_t q = P; q > 1; q >>= 1) --lo;
        while (lo > 11) { P <<= 1; --lo; }                       // <= 2048 buckets per partition (shared-memory histogram)
        while ((size_t)p.W * P > 8192 && lo < 11) { P >>= 1; ++lo; }   // <= 32 KB of partition counters per k_digits block
        if ((size_t)p.W * P <= 8192 && P >= 1 && ((uint32_t)P << lo) == p.NB) {
            p.sort_P = P;
            p.sort_lo_bits = lo;
        }
    }
    return p;
}

// ---------------------------------------------------------------------------------------------------------------------
// input preparation / validation
// ---------------------------------------------------------------------------------------------------------------------
// `endo` != 0: out has room for 2 n points and also receives phi(P_i) = (beta x_i, y_i) at index n + i (glv.cuh).
__global__ void k_points_prepare(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, size_t n, int format, int check, int endo,
                                 int* __restrict__ status) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        G1Affine p = g1_affine_load(in, i);
        // raw limbs must be reduced in EITHER format: zero tests and equality tests downstream assume a unique representative
        if (check && (!fp_is_canonical(p.x) || !fp_is_canonical(p.y))) atomicCAS(status, 0, SNARKV_ERR_BAD_POINT);
        if (format == SNARKV_CANONICAL) {
            p.x = fp_to_mont(p.x);
            p.y = fp_to_mont(p.y);
        }
        if (check && !g1_affine_is_on_curve(p)) atomicCAS(status, 0, SNARKV_ERR_BAD_POINT);
        if (out) g1_affine_store(out, i, p);
        if (endo) {
            constexpr uint32_t beta_limbs[8] = SNARKV_GLV_BETA_MONT_LIMBS;
            Fq beta;
#pragma unroll
            for (int j = 0; j < 8; ++j) beta.v[j] = beta_limbs[j];
            p.x = fp_mul(p.x, beta);          // the identity (0, 0) maps to (0, 0)
            g1_affine_store(out, n + i, p);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// K1 / K3: signed-digit decomposition.  `windowed_scalar` of util/msm.rs:271-281 extracts unsigned c-bit digits from the
// canonical little-endian repr; here each digit d in [0, 2^c) plus the carry from below is mapped to (-2^(c-1), 2^(c-1)].
// ---------------------------------------------------------------------------------------------------------------------
// ---- TMA (bulk async copy) helpers: 1-D cp.async.bulk global -> shared with mbarrier completion (SASS: UBLKCP / SYNCS) ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}

// K1: every scalar -> W signed digits, stored window-major (digits[w][i] = |d| | sign << 31, coalesced) + histogram.
// The scalar stream is staged through shared memory by TMA bulk copies: one elected thread issues an 8 KB
// cp.async.bulk per 256-scalar tile into a double buffer while the block decomposes the previous tile.
#define SNARKV_DIGIT_TILE 256
// PART (large inputs, sort.cuh): no global histogram — `counters` is the W x P array of COARSE partition sizes, accumulated in
// (dynamic) shared memory and flushed once per block; the bucket sizes proper come out of k_sort_buckets.
template <bool GLV, bool PART>
__global__ void __launch_bounds__(SNARKV_DIGIT_TILE) k_digits(const uint8_t* __restrict__ scalars, size_t n, int format, int check, uint32_t c,
                                                              uint32_t W, uint32_t NB, uint32_t* __restrict__ counters,
                                                              uint32_t* __restrict__ digits, int* __restrict__ status, uint32_t P,
                                                              uint32_t lo_bits) {
    __shared__ alignas(128) uint8_t tile[2][SNARKV_DIGIT_TILE * 32];
    __shared__ alignas(8) uint64_t bar[2];
    extern __shared__ uint32_t part_hist[];   // PART: W x P
    if (PART) {
        for (uint32_t k = threadIdx.x; k < W * P; k += SNARKV_DIGIT_TILE) part_hist[k] = 0;
    }
    auto count = [&](uint32_t w, uint32_t d) {
        if (PART) atomicAdd(&part_hist[w * P + ((d - 1u) >> lo_bits)], 1u);
        else atomicAdd(&counters[w * NB + (d - 1u)], 1u);
    };
    const uint32_t mask = (1u << c) - 1u;
    const uint32_t tid = threadIdx.x;
    const size_t ntiles = (n + SNARKV_DIGIT_TILE - 1) / SNARKV_DIGIT_TILE;
    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](size_t t, uint32_t buf) {
        const size_t first = t * SNARKV_DIGIT_TILE;
        const uint32_t cnt = (uint32_t)((n - first < SNARKV_DIGIT_TILE) ? n - first : SNARKV_DIGIT_TILE);
        tma_load_1d(tile[buf], scalars + first * 32, cnt * 32u, &bar[buf]);
    };
    size_t t = blockIdx.x;
    if (tid == 0 && t < ntiles) issue(t, 0);
    for (uint32_t it = 0; t < ntiles; t += gridDim.x, ++it) {
        const uint32_t buf = it & 1u;
        const size_t nxt = t + gridDim.x;
        if (tid == 0 && nxt < ntiles) issue(nxt, buf ^ 1u);   // buf^1 was released by the __syncthreads closing iteration it-1
        mbar_wait(&bar[buf], (it >> 1) & 1u);
        const size_t i = t * SNARKV_DIGIT_TILE + tid;
        if (i < n) {
            Fr s;
            {
                const uint4* q = reinterpret_cast<const uint4*>(&tile[buf][tid * 32]);
                uint4 lo = q[0], hi = q[1];
                s.v[0] = lo.x; s.v[1] = lo.y; s.v[2] = lo.z; s.v[3] = lo.w; s.v[4] = hi.x; s.v[5] = hi.y; s.v[6] = hi.z; s.v[7] = hi.w;
            }
            if (check && !fp_is_canonical(s)) atomicCAS(status, 0, SNARKV_ERR_BAD_SCALAR);   // raw limbs >= r in either format
            if (format == SNARKV_MONTGOMERY) s = fp_from_mont(s);
            if (!GLV) {
                uint32_t carry = 0;
                for (uint32_t w = 0; w < W; ++w) {
                    uint32_t d = (s.v[0] & mask) + carry;
                    // s >>= c  (c < 32)
#pragma unroll
                    for (int j = 0; j < 7; ++j) s.v[j] = __funnelshift_r(s.v[j], s.v[j + 1], c);
                    s.v[7] >>= c;
                    uint32_t neg = 0;
                    if (d > NB) {
                        d = (mask + 1u) - d;
                        neg = 1u;
                        carry = 1u;
                    } else carry = 0u;
                    digits[(size_t)w * n + i] = d | (neg << 31);
                    if (d != 0) count(w, d);
                }
            } else {
                // two half-length virtual terms: index i carries |k1| (sign neg1) for P_i, index n + i carries |k2| for phi(P_i)
                uint32_t mag[2][5], sgn[2];
                glv::decompose(s.v, mag[0], sgn[0], mag[1], sgn[1]);
                const size_t nv = 2 * n;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint32_t m0 = mag[h][0], m1 = mag[h][1], m2 = mag[h][2], m3 = mag[h][3], m4 = mag[h][4];
                    uint32_t carry = 0;
                    for (uint32_t w = 0; w < W; ++w) {
                        uint32_t d = (m0 & mask) + carry;
                        m0 = __funnelshift_r(m0, m1, c); m1 = __funnelshift_r(m1, m2, c); m2 = __funnelshift_r(m2, m3, c);
                        m3 = __funnelshift_r(m3, m4, c); m4 >>= c;
                        uint32_t neg = 0;
                        if (d > NB) {
                            d = (mask + 1u) - d;
                            neg = 1u;
                            carry = 1u;
                        } else carry = 0u;
                        digits[(size_t)w * nv + i + (size_t)h * n] = d | ((neg ^ sgn[h]) << 31);
                        if (d != 0) count(w, d);
                    }
                }
            }
        }
        __syncthreads();
    }
    if (PART) {
        for (uint32_t k = threadIdx.x; k < W * P; k += SNARKV_DIGIT_TILE) {
            const uint32_t v = part_hist[k];
            if (v) atomicAdd(&counters[k], v);
        }
    }
}

// K3: window-major scatter.  Consecutive blocks work on the same window, so the random 4-byte stores of a window land in a
// n x 4 B region (64 MB at 2^24 terms) that stays resident in the 126 MB L2 until its 32-byte sectors are complete.
__global__ void __launch_bounds__(256) k_scatter(const uint32_t* __restrict__ digits, size_t n, uint32_t W, uint32_t NB,
                                                 uint32_t* __restrict__ cursor, uint32_t* __restrict__ sorted) {
    // blockIdx.y = window (blocks are dispatched x-fastest, so neighbouring blocks share a window); no per-element division
    const uint32_t w = blockIdx.y;
    const uint32_t* dg = digits + (size_t)w * n;
    uint32_t* cur = cursor + (size_t)w * NB;
    uint32_t* out = sorted + (size_t)w * n;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t e = dg[i];
        const uint32_t d = e & 0x7fffffffu;
        if (d == 0) continue;
        const uint32_t pos = atomicAdd(&cur[d - 1u], 1u);
        out[pos] = (uint32_t)i | (e & 0x80000000u);
    }
}

// K2: per-window exclusive scans, one block per window: `offsets` = start of each bucket's run in the sorted array,
// `task_base` = index of the bucket's first accumulate task (a bucket of cnt points owns ceil(cnt / T) tasks).
// The window's NB counters are walked in tiles of blockDim.x consecutive entries (coalesced loads and stores); each tile is
// scanned with warp shuffles + one shared-memory hop and chained through a running carry.
__global__ void __launch_bounds__(1024) k_scan(const uint32_t* __restrict__ counts, uint32_t* __restrict__ offsets,
                                               uint32_t* __restrict__ cursor, uint32_t* __restrict__ task_base,
                                               uint32_t* __restrict__ window_tasks, uint32_t NB, uint32_t T) {
    __shared__ uint2 warp_tot[32];
    __shared__ uint2 carry_sm;
    const uint32_t w = blockIdx.x, t = threadIdx.x, NT = blockDim.x;
    const uint32_t lane = t & 31, wid = t >> 5;
    const size_t base = (size_t)w * NB;
    if (t == 0) carry_sm = make_uint2(0, 0);
    __syncthreads();
    for (uint32_t tile = 0; tile < NB; tile += NT) {
        const uint32_t k = tile + t;
        const uint32_t c = k < NB ? counts[base + k] : 0u;
        const uint2 v = make_uint2(c, (c + T - 1) / T);
        uint2 incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t vx = __shfl_up_sync(0xffffffffu, incl.x, o), vy = __shfl_up_sync(0xffffffffu, incl.y, o);
            if (lane >= (uint32_t)o) { incl.x += vx; incl.y += vy; }
        }
        if (lane == 31) warp_tot[wid] = incl;
        // read the running totals BEFORE the barrier: warp 0 overwrites them right after it (racecheck: a warp still reading
        // carry_sm there could see the next tile's value)
        const uint2 carry = carry_sm;
        __syncthreads();
        if (wid == 0) {
            uint2 x = lane < (NT >> 5) ? warp_tot[lane] : make_uint2(0, 0);
            uint2 ix = x;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t ux = __shfl_up_sync(0xffffffffu, ix.x, o), uy = __shfl_up_sync(0xffffffffu, ix.y, o);
                if (lane >= (uint32_t)o) { ix.x += ux; ix.y += uy; }
            }
            warp_tot[lane] = make_uint2(ix.x - x.x, ix.y - x.y);                        // exclusive warp offsets
            if (lane == 31) carry_sm = make_uint2(carry.x + ix.x, carry.y + ix.y);   // running totals for the next tile
        }
        __syncthreads();
        if (k < NB) {
            const uint32_t off = carry.x + warp_tot[wid].x + incl.x - v.x;
            offsets[base + k] = off;
            cursor[base + k] = off;
            task_base[base + k] = carry.y + warp_tot[wid].y + incl.y - v.y;
        }
        __syncthreads();
    }
    if (t == 0) window_tasks[w] = carry_sm.y;
}

// K2b: task descriptors.  Task slot (w, task_base[b] + j) = (bucket b, part j).
__global__ void __launch_bounds__(256) k_build_tasks(const uint32_t* __restrict__ counts, const uint32_t* __restrict__ task_base,
                                                     uint32_t NB, uint32_t total, uint32_t T, uint32_t cap, uint2* __restrict__ tasks) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= total) return;
    const uint32_t w = tid / NB, b = tid - w * NB;
    const uint32_t parts = (counts[tid] + T - 1) / T;
    uint2* out = tasks + (size_t)w * cap + task_base[tid];
    for (uint32_t j = 0; j < parts; ++j) out[j] = make_uint2(b, j);
}

// K2c: order each window's tasks by length (longest first) with a shared-memory counting sort over 1024 quantised length
// bins, so that the 32 tasks of a warp run for (almost) the same number of additions and the long tasks start first.
#define SNARKV_ORDER_BINS 1024
__global__ void __launch_bounds__(1024) k_order_tasks(const uint32_t* __restrict__ counts, const uint2* __restrict__ tasks,
                                                      const uint32_t* __restrict__ window_tasks, uint32_t NB, uint32_t T, uint32_t cap,
                                                      uint32_t* __restrict__ order) {
    __shared__ uint32_t hist[SNARKV_ORDER_BINS];
    __shared__ uint32_t warp_tot[32];
    const uint32_t w = blockIdx.x, t = threadIdx.x;
    const uint32_t nt = window_tasks[w];
    const uint2* tw = tasks + (size_t)w * cap;
    hist[t] = 0;
    __syncthreads();
    auto bin_of = [&](uint2 task) -> uint32_t {
        const uint32_t len = min(T, counts[(size_t)w * NB + task.x] - task.y * T);       // 1..T
        const uint32_t q = (uint32_t)(((uint64_t)len * (SNARKV_ORDER_BINS - 1)) / T);    // 0..BINS-1
        return (SNARKV_ORDER_BINS - 1) - q;                                              // longest first
    };
    for (uint32_t i = t; i < nt; i += blockDim.x) atomicAdd(&hist[bin_of(tw[i])], 1u);
    __syncthreads();
    // exclusive scan of the 1024 bins (one per thread)
    const uint32_t v = hist[t];
    uint32_t incl = v;
    const uint32_t lane = t & 31, wid = t >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (uint32_t)o) incl += u;
    }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        uint32_t x = warp_tot[lane], ix = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t u = __shfl_up_sync(0xffffffffu, ix, o);
            if (lane >= (uint32_t)o) ix += u;
        }
        warp_tot[lane] = ix - x;
    }
    __syncthreads();
    hist[t] = warp_tot[wid] + incl - v;
    __syncthreads();
    uint32_t* ow = order + (size_t)w * cap;
    for (uint32_t i = t; i < nt; i += blockDim.x) ow[atomicAdd(&hist[bin_of(tw[i])], 1u)] = i;
}

// ---------------------------------------------------------------------------------------------------------------------
// K4: bucket accumulation — the multi-scalar multiplication proper.  `buckets[scalar - 1].add_assign(base)` of
// util/msm.rs:291-296, with Bucket::{None, Affine, Projective} (util/msm.rs:228-246) collapsed into the XYZZ identity test.
// One thread per (window, bucket); the next point is fetched (4 x 128-bit loads) while the current addition runs.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_bucket_accumulate(const uint8_t* __restrict__ points0, const uint8_t* __restrict__ points1,
                                                           const uint32_t* __restrict__ sorted,
                                                           const uint32_t* __restrict__ offsets, const uint32_t* __restrict__ counts,
                                                           const uint2* __restrict__ tasks, const uint32_t* __restrict__ window_tasks,
                                                           const uint32_t* __restrict__ order, size_t n, uint32_t NB, uint32_t T,
                                                           uint32_t cap, uint8_t* __restrict__ task_out) {
    const uint32_t w = blockIdx.y;
    const uint32_t rank = blockIdx.x * blockDim.x + threadIdx.x;
    if (rank >= window_tasks[w]) return;
    // blockIdx.z selects one of the base sets that share these scalars (KzgAs::verify: lhs and rhs points, accumulation.rs:53-60)
    const uint8_t* __restrict__ points = blockIdx.z == 0 ? points0 : points1;
    task_out += (size_t)blockIdx.z * gridDim.y * cap * 128;
    const uint32_t slot = order[(size_t)w * cap + rank];
    const uint2 task = tasks[(size_t)w * cap + slot];
    const uint32_t bucket = w * NB + task.x;
    const uint32_t first = task.y * T;
    const uint32_t cnt = min(T, counts[bucket] - first);
    const uint32_t* list = sorted + (size_t)w * n + offsets[bucket] + first;
    G1Xyzz acc = xyzz_identity();
    uint32_t e = list[0];
    G1Affine nxt = g1_affine_load(points, e & 0x7fffffffu);
#pragma unroll 1
    for (uint32_t k = 0; k < cnt; ++k) {
        G1Affine cur = nxt;
        const uint32_t neg = e >> 31;
        if (k + 1 < cnt) {
            e = list[k + 1];
            nxt = g1_affine_load(points, e & 0x7fffffffu);
        }
        if (g1_affine_is_identity(cur)) continue;
        if (neg) cur.y = fp_neg(cur.y);
        xyzz_madd(acc, cur.x, cur.y);
    }
    xyzz_store(task_out, (size_t)w * cap + slot, acc);
}

// K4b: bucket = sum of its task results (one copy in the common case of one task per bucket).  Buckets split into more
// than MERGE_SERIAL tasks (heavily skewed digit distributions) are queued for the block-wide merge below.
#define SNARKV_MERGE_SERIAL 16u
__global__ void __launch_bounds__(128) k_bucket_merge(const uint8_t* __restrict__ task_out, const uint32_t* __restrict__ counts,
                                                      const uint32_t* __restrict__ task_base, uint32_t NB, uint32_t total, uint32_t T,
                                                      uint32_t cap, uint8_t* __restrict__ buckets, uint32_t* __restrict__ big_count,
                                                      uint32_t* __restrict__ big_list, int add_existing) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= total) return;
    const uint32_t w = tid / NB;
    const uint32_t parts = (counts[tid] + T - 1) / T;
    task_out += (size_t)blockIdx.y * (total / NB) * cap * 128;
    buckets += (size_t)blockIdx.y * total * 128;
    if (parts > SNARKV_MERGE_SERIAL) {
        if (blockIdx.y == 0) big_list[atomicAdd(big_count, 1u)] = tid;   // the list depends on the shared counts only
        return;
    }
    const size_t base = (size_t)w * cap + task_base[tid];
    // add_existing: this launch belongs to a later term-chunk of the same MSM; the bucket already holds the earlier chunks' sum
    if (add_existing && parts == 0) return;
    G1Xyzz acc = add_existing ? xyzz_load(buckets, tid) : xyzz_identity();
    if (parts == 1 && !add_existing) acc = xyzz_load(task_out, base);
    else
        for (uint32_t j = 0; j < parts; ++j) acc = xyzz_add(acc, xyzz_load(task_out, base + j));
    xyzz_store(buckets, tid, acc);
}
__global__ void __launch_bounds__(128) k_bucket_merge_big(const uint8_t* __restrict__ task_out, const uint32_t* __restrict__ counts,
                                                          const uint32_t* __restrict__ task_base, uint32_t NB, uint32_t T, uint32_t cap,
                                                          uint8_t* __restrict__ buckets, const uint32_t* __restrict__ big_count,
                                                          const uint32_t* __restrict__ big_list, uint32_t W, int add_existing) {
    __shared__ G1Xyzz sm[128];
    const uint32_t t = threadIdx.x;
    task_out += (size_t)blockIdx.y * W * cap * 128;
    buckets += (size_t)blockIdx.y * W * NB * 128;
    for (uint32_t q = blockIdx.x; q < *big_count; q += gridDim.x) {
        const uint32_t tid = big_list[q];
        const uint32_t w = tid / NB;
        const uint32_t parts = (counts[tid] + T - 1) / T;
        const size_t base = (size_t)w * cap + task_base[tid];
        G1Xyzz acc = xyzz_identity();
        for (uint32_t j = t; j < parts; j += blockDim.x) acc = xyzz_add(acc, xyzz_load(task_out, base + j));
        sm[t] = acc;
        __syncthreads();
        for (uint32_t s = blockDim.x >> 1; s > 0; s >>= 1) {
            if (t < s) sm[t] = xyzz_add(sm[t], sm[t + s]);
            __syncthreads();
        }
        if (t == 0) xyzz_store(buckets, tid, add_existing ? xyzz_add(xyzz_load(buckets, tid), sm[0]) : sm[0]);
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// K5: per-window bucket reduction  sum_{b=1..NB} b * B_b — the running-sum loop of util/msm.rs:298-302, cut into
// segments of `seg` buckets: a segment covering array indices [lo, hi) contributes
//   sum (idx - lo + 1) B_idx  (running sum)  +  lo * sum B_idx  (small scalar multiple).
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_bucket_reduce(const uint8_t* __restrict__ buckets, uint32_t NB, uint32_t seg, uint32_t J,
                                                       uint8_t* __restrict__ segpart) {
    __shared__ G1Xyzz sm[128];
    const uint32_t w = blockIdx.y, t = threadIdx.x;
    const uint32_t j = blockIdx.x * blockDim.x + t;
    buckets += (size_t)blockIdx.z * gridDim.y * NB * 128;
    segpart += (size_t)blockIdx.z * gridDim.y * gridDim.x * 128;
    G1Xyzz acc = xyzz_identity();
    if (j < J) {
        const uint32_t lo = j * seg, hi = min(lo + seg, NB);
        G1Xyzz running = xyzz_identity();
        for (uint32_t idx = hi; idx-- > lo;) {
            G1Xyzz b = xyzz_load(buckets, (size_t)w * NB + idx);
            running = xyzz_add(running, b);
            acc = xyzz_add(acc, running);
        }
        if (lo > 0) acc = xyzz_add(acc, xyzz_mul_small(running, lo));
    }
    // the block's 128 segment results are summed here (7 parallel levels): the per-window kernel that follows has 128x fewer
    // values to fold (measured: c = 16 0.45 + 0.18 -> 0.52 + 0.05 ms; c = 17 with seg 32: 1.08 + 0.27 -> 0.89 + 0.05 ms).
    // Measured and rejected: replacing the per-thread small multiple lo * running by a weighted (T, S) tree — its doublings are
    // serial inside the block (0.77 + 0.17 ms at c = 16).
    sm[t] = acc;
    __syncthreads();
    for (uint32_t s = blockDim.x >> 1; s > 0; s >>= 1) {
        if (t < s) sm[t] = xyzz_add(sm[t], sm[t + s]);
        __syncthreads();
    }
    if (t == 0) xyzz_store(segpart, (size_t)w * gridDim.x + blockIdx.x, sm[0]);
}

// K6a: sum the J segment results of one window; one block per window, shared-memory tree.
__global__ void __launch_bounds__(128) k_window_sum(const uint8_t* __restrict__ segpart, uint32_t J, uint8_t* __restrict__ winsum) {
    __shared__ G1Xyzz sm[128];
    const uint32_t w = blockIdx.x, t = threadIdx.x;
    segpart += (size_t)blockIdx.y * gridDim.x * J * 128;
    winsum += (size_t)blockIdx.y * gridDim.x * 128;
    G1Xyzz acc = xyzz_identity();
    for (uint32_t j = t; j < J; j += blockDim.x) acc = xyzz_add(acc, xyzz_load(segpart, (size_t)w * J + j));
    sm[t] = acc;
    __syncthreads();
    for (uint32_t s = blockDim.x >> 1; s > 0; s >>= 1) {
        if (t < s) sm[t] = xyzz_add(sm[t], sm[t + s]);
        __syncthreads();
    }
    if (t == 0) xyzz_store(winsum, w, sm[0]);
}

__device__ __forceinline__ void store_affine_fmt(void* out, const G1Affine& a, int format) {
    G1Affine o = a;
    if (format == SNARKV_CANONICAL) {
        o.x = fp_from_mont(o.x);
        o.y = fp_from_mont(o.y);
    }
    g1_affine_store(out, 0, o);
}
__device__ __forceinline__ void store_jacobian(void* out, const G1Jac& j) {
    uint8_t* p = reinterpret_cast<uint8_t*>(out);
    fp_store<FQ>(p, j.x);
    fp_store<FQ>(p + 32, j.y);
    fp_store<FQ>(p + 64, j.z);
}

// K6b: result = sum_w 2^(c w) * winsum[w]  (the `result.double()` x window_size loop of util/msm.rs:285-287), then the
// caller's `.to_affine()` (native.rs:70).
__global__ void k_msm_final(const uint8_t* __restrict__ winsum, uint32_t W, uint32_t c, int format, void* out_affine,
                            void* out_jacobian) {
    // one warp (= one block) per base set; every aligned group of 4 lanes runs the same 4-lane point arithmetic
    // (g1.cuh: *_x4), thread 0 stores
    __shared__ Fq xch[4];
    winsum += (size_t)blockIdx.x * W * 128;
    if (out_affine) out_affine = (uint8_t*)out_affine + 64 * blockIdx.x;
    if (out_jacobian) out_jacobian = (uint8_t*)out_jacobian + 96 * blockIdx.x;
    const int lane = threadIdx.x & 3;
    G1Xyzz acc = xyzz_load(winsum, W - 1);
    for (uint32_t w = W - 1; w-- > 0;) {
        for (uint32_t k = 0; k < c; ++k) acc = xyzz_dbl_x4(acc, lane, xch);
        acc = xyzz_add_x4(acc, xyzz_load(winsum, w), lane, xch);
    }
    if (out_jacobian && threadIdx.x == 0) store_jacobian(out_jacobian, xyzz_to_jacobian(acc));
    if (out_affine) {
        G1Affine a = xyzz_to_affine_serial(acc);
        if (threadIdx.x == 0) store_affine_fmt(out_affine, a, format);
    }
}

// fold of per-GPU Jacobian partials (util/msm.rs:333-335) + to_affine
__global__ void k_fold_partials(const uint8_t* __restrict__ partials, uint32_t k, int format, void* out_affine, void* out_jacobian) {
    __shared__ Fq xch[4];
    const int lane = threadIdx.x & 3;
    G1Xyzz acc = xyzz_identity();
    for (uint32_t i = 0; i < k; ++i) {
        G1Jac j;
        const uint8_t* p = partials + (size_t)i * 96;
        j.x = fp_load<FQ>(p); j.y = fp_load<FQ>(p + 32); j.z = fp_load<FQ>(p + 64);
        acc = xyzz_add_x4(acc, jacobian_to_xyzz(j), lane, xch);
    }
    if (out_jacobian && threadIdx.x == 0) store_jacobian(out_jacobian, xyzz_to_jacobian(acc));
    if (out_affine) {
        G1Affine a = xyzz_to_affine_serial(acc);
        if (threadIdx.x == 0) store_affine_fmt(out_affine, a, format);
    }
}

// The same fold fused with the gather: partial i is read straight out of the memory of the GPU that produced it (peer access over
// NVLink: `src.p[i]` points into device i's HBM), so the one exchange step of the chunk-partitioned MSM (util/msm.rs:333-335 across
// GPUs) needs no separate collective or staging copy — 96 bytes per peer cross the link inside this kernel.
struct PartialPtrs { const uint8_t* p[SNARKV_MAX_DEVICES]; };
__global__ void k_fold_partials_peer(PartialPtrs src, uint32_t k, int format, void* out_affine, void* out_jacobian) {
    __shared__ Fq xch[4];
    const int lane = threadIdx.x & 3;
    G1Xyzz acc = xyzz_identity();
    for (uint32_t i = 0; i < k; ++i) {
        G1Jac j;
        const uint8_t* p = src.p[i];
        j.x = fp_load_rw<FQ>(p); j.y = fp_load_rw<FQ>(p + 32); j.z = fp_load_rw<FQ>(p + 64);   // coherent loads: peer memory
        acc = xyzz_add_x4(acc, jacobian_to_xyzz(j), lane, xch);
    }
    if (out_jacobian && threadIdx.x == 0) store_jacobian(out_jacobian, xyzz_to_jacobian(acc));
    if (out_affine) {
        G1Affine a = xyzz_to_affine_serial(acc);
        if (threadIdx.x == 0) store_affine_fmt(out_affine, a, format);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// host orchestration
// ---------------------------------------------------------------------------------------------------------------------
// Phase 1 needs only the scalars (K1-K3); phase 2 needs the points (K4-K6).  The host-buffer entry point issues the
// points' host->device copy between the two so that it overlaps the sort.
struct MsmWork {
    MsmPlan pl;
    int* status;
    uint32_t *counts, *offsets, *cursor, *sorted, *digits;
    uint32_t *part_count, *part_off, *part_cursor, *rec_idx;   // two-level sort (sort.cuh)
    uint16_t* rec_lo;
    uint32_t *task_base, *window_tasks, *big;   // big = [count | list of bucket ids]
    uint2* tasks;
    uint32_t* order;
    uint8_t *task_out, *buckets, *segpart, *winsum;
};

static int msm_alloc(snarkv_ctx* ctx, size_t n, void* d_status, MsmWork& wk, int B = 1, int glv_mode = -1) {
    if (n == 0) return ctx->fail(SNARKV_ERR_EMPTY, "multi_scalar_multiplication on an empty slice");
    if (n >= (1ull << 31)) return ctx->fail(SNARKV_ERR_USAGE, "n must be < 2^31");
    wk.pl = make_plan(n, ctx->window_bits, glv_mode >= 0 ? glv_mode : ctx->glv_mode);
    const MsmPlan& pl = wk.pl;
    // sorted references keep the (virtual) term index in 31 bits next to the sign bit
    if (plan_virtual_terms(pl, n) >= (1ull << 31)) return ctx->fail(SNARKV_ERR_USAGE, "GLV doubles the term count: n must be < 2^30 with snarkv_set_glv_mode(1)");
    const size_t nbk = (size_t)pl.W * pl.NB;
    wk.status = (int*)d_status;
    if (!wk.status) wk.status = (int*)ctx->wsget(WS_STATUS, 4);
    wk.counts = (uint32_t*)ctx->wsget(WS_COUNTS, nbk * 4);
    wk.offsets = (uint32_t*)ctx->wsget(WS_OFFSETS, nbk * 4);
    wk.cursor = (uint32_t*)ctx->wsget(WS_CURSOR, nbk * 4);
    const size_t nv = plan_virtual_terms(pl, n);
    wk.sorted = (uint32_t*)ctx->wsget(WS_SORTED, (size_t)pl.W * nv * 4);
    wk.digits = (uint32_t*)ctx->wsget(WS_DIGITS, (size_t)pl.W * nv * 4);
    wk.part_count = wk.part_off = wk.part_cursor = wk.rec_idx = nullptr;
    wk.rec_lo = nullptr;
    if (pl.sort_P) {
        uint32_t* pc = (uint32_t*)ctx->wsget(WS_SORT_PART, (size_t)3 * pl.W * pl.sort_P * 4);
        wk.rec_idx = (uint32_t*)ctx->wsget(WS_SORT_REC_IDX, (size_t)pl.W * nv * 4 + 64);   // + slack: k_sort_buckets reads aligned vectors
        wk.rec_lo = (uint16_t*)ctx->wsget(WS_SORT_REC_LO, (size_t)pl.W * nv * 2 + 64);
        if (!pc || !wk.rec_idx || !wk.rec_lo) return SNARKV_ERR_CUDA;
        wk.part_count = pc;
        wk.part_off = pc + (size_t)pl.W * pl.sort_P;
        wk.part_cursor = pc + (size_t)2 * pl.W * pl.sort_P;
    }
    wk.buckets = (uint8_t*)ctx->wsget(WS_BUCKETS, (size_t)B * nbk * 128);
    wk.segpart = (uint8_t*)ctx->wsget(WS_SEGPART, (size_t)B * pl.W * pl.J * 128);
    wk.winsum = (uint8_t*)ctx->wsget(WS_WINSUM, (size_t)B * pl.W * 128);
    wk.task_base = (uint32_t*)ctx->wsget(WS_TASK_BASE, nbk * 4);
    wk.window_tasks = (uint32_t*)ctx->wsget(WS_WINDOW_TASKS, (size_t)pl.W * 4);
    wk.big = (uint32_t*)ctx->wsget(WS_BIG, (nbk + 1) * 4);
    wk.tasks = (uint2*)ctx->wsget(WS_TASKS, (size_t)pl.W * pl.cap * 8);
    wk.task_out = (uint8_t*)ctx->wsget(WS_TASK_OUT, (size_t)B * pl.W * pl.cap * 128);
    wk.order = (uint32_t*)ctx->wsget(WS_ORDER, (size_t)pl.W * pl.cap * 4);
    if (!wk.status || !wk.counts || !wk.offsets || !wk.cursor || !wk.sorted || !wk.buckets || !wk.segpart || !wk.winsum ||
        !wk.digits || !wk.task_base || !wk.window_tasks || !wk.big || !wk.tasks || !wk.task_out || !wk.order)
        return SNARKV_ERR_CUDA;
    return SNARKV_OK;
}

static int msm_sort_phase(snarkv_ctx* ctx, const MsmWork& wk, const void* d_scalars, size_t n, int scalar_format, int check) {
    const MsmPlan& pl = wk.pl;
    cudaStream_t st = ctx->stream;
    const size_t nbk = (size_t)pl.W * pl.NB;
    const size_t want = (n + SNARKV_DIGIT_TILE - 1) / SNARKV_DIGIT_TILE, cap = (size_t)ctx->sm_count * 8;
    const int dig_blocks = (int)(want < cap ? want : cap);
    if ((reinterpret_cast<uintptr_t>(d_scalars) & 15u) != 0) return ctx->fail(SNARKV_ERR_USAGE, "scalar buffer must be 16-byte aligned");
    const size_t nvt = plan_virtual_terms(pl, n);
    if (pl.sort_P) {
        // large inputs: two-level partition without per-element global atomics (sort.cuh)
        const uint32_t P = pl.sort_P, lo = pl.sort_lo_bits;
        {
            Stage sg(ctx, "msm_digits");
            SNARKV_CUDA_TRY(ctx, cudaMemsetAsync(wk.status, 0, 4, st));
            SNARKV_CUDA_TRY(ctx, cudaMemsetAsync(wk.part_count, 0, (size_t)pl.W * P * 4, st));
            const size_t hist_bytes = (size_t)pl.W * P * 4;
            auto kd = pl.glv ? k_digits<true, true> : k_digits<false, true>;
            SNARKV_CUDA_TRY(ctx, cudaFuncSetAttribute(kd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hist_bytes));
            kd<<<dig_blocks, 256, hist_bytes, st>>>((const uint8_t*)d_scalars, n, scalar_format, check, pl.c, pl.W, pl.NB, wk.part_count, wk.digits,
                                                    wk.status, P, lo);
            SNARKV_LAUNCH_CHECK(ctx, "k_digits<part>");
            sg.launched();
        }
        {
            Stage sg(ctx, "msm_sort_partition");
            k_part_scan<<<pl.W < 64 ? pl.W : 64, 1024, 0, st>>>(wk.part_count, wk.part_off, wk.part_cursor, pl.W, P);
            SNARKV_LAUNCH_CHECK(ctx, "k_part_scan");
            sg.launched();
            const int tile = ctx->sort_tile;                    // digits per work item: 4096 (256 threads) or 8192 (512 threads)
            const size_t smem = (size_t)3 * P * 4 + (size_t)tile * 8;
            auto kp = tile == 8192 ? k_partition<512> : k_partition<256>;
            SNARKV_CUDA_TRY(ctx, cudaFuncSetAttribute(kp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            const size_t items = ((nvt + tile - 1) / tile) * pl.W;
            const size_t want_b = items, cap_b = (size_t)ctx->sm_count * (ctx->sort_blocks_per_sm > 0 ? ctx->sort_blocks_per_sm : 4);
            kp<<<(unsigned)(want_b < cap_b ? want_b : cap_b), tile / SNARKV_SORT_PER_THREAD, smem, st>>>(wk.digits, nvt, pl.W, P, lo, wk.part_off,
                                                                                                   wk.part_cursor, wk.rec_idx, wk.rec_lo);
            SNARKV_LAUNCH_CHECK(ctx, "k_partition");
            sg.launched();
        }
        {
            Stage sg(ctx, "msm_sort_buckets");
            const size_t smem = ((size_t)2 * (1u << lo) + 1 + SNARKV_SORT_CAP) * 4;
            SNARKV_CUDA_TRY(ctx, cudaFuncSetAttribute(k_sort_buckets, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_sort_buckets<<<dim3(P, pl.W), SNARKV_SORT_BUCKET_THREADS, smem, st>>>(wk.rec_idx, wk.rec_lo, nvt, P, lo, pl.NB, wk.part_off, wk.part_count,
                                                                                   wk.counts, wk.sorted);
            SNARKV_LAUNCH_CHECK(ctx, "k_sort_buckets");
            sg.launched();
        }
    } else {
        Stage sg(ctx, "msm_digits_count");
        SNARKV_CUDA_TRY(ctx, cudaMemsetAsync(wk.status, 0, 4, st));
        SNARKV_CUDA_TRY(ctx, cudaMemsetAsync(wk.counts, 0, nbk * 4, st));
        if (pl.glv)
            k_digits<true, false><<<dig_blocks, 256, 0, st>>>((const uint8_t*)d_scalars, n, scalar_format, check, pl.c, pl.W, pl.NB, wk.counts,
                                                              wk.digits, wk.status, 0, 0);
        else
            k_digits<false, false><<<dig_blocks, 256, 0, st>>>((const uint8_t*)d_scalars, n, scalar_format, check, pl.c, pl.W, pl.NB, wk.counts,
                                                               wk.digits, wk.status, 0, 0);
        SNARKV_LAUNCH_CHECK(ctx, "k_digits");
        sg.launched();
    }
    {
        Stage sg(ctx, "msm_scan");
        k_scan<<<pl.W, 1024, 0, st>>>(wk.counts, wk.offsets, wk.cursor, wk.task_base, wk.window_tasks, pl.NB, pl.T);
        SNARKV_LAUNCH_CHECK(ctx, "k_scan");
        sg.launched();
        k_build_tasks<<<(unsigned)((nbk + 255) / 256), 256, 0, st>>>(wk.counts, wk.task_base, pl.NB, (uint32_t)nbk, pl.T, pl.cap, wk.tasks);
        SNARKV_LAUNCH_CHECK(ctx, "k_build_tasks");
        sg.launched();
        k_order_tasks<<<pl.W, 1024, 0, st>>>(wk.counts, wk.tasks, wk.window_tasks, pl.NB, pl.T, pl.cap, wk.order);
        SNARKV_LAUNCH_CHECK(ctx, "k_order_tasks");
        sg.launched();
    }
    if (!pl.sort_P) {
        Stage sg(ctx, "msm_digits_scatter");
        const size_t nv = plan_virtual_terms(pl, n);
        // 8 elements per thread; at most 65535 blocks per window
        size_t gx = (nv + 2047) / 2048;
        if (gx > 65535) gx = 65535;
        k_scatter<<<dim3((unsigned)gx, pl.W), 256, 0, st>>>(wk.digits, nv, pl.W, pl.NB, wk.cursor, wk.sorted);
        SNARKV_LAUNCH_CHECK(ctx, "k_scatter");
        sg.launched();
    }
    return SNARKV_OK;
}

// Point phase, first half: (convert,) gather-accumulate the sorted terms into task results and merge them into the bucket
// array.  B = 1 or 2 base sets share the scalars / sort of `wk`.  `add_existing`: the buckets already hold the sums of earlier
// term-chunks of the same MSM (host entry point pipeline) and this chunk is added on top.  `conv_dst`: caller-provided
// destination for the Montgomery copy of CANONICAL points (B must be 1).
// `prepared`: d_points is a resident base set (snarkv_g1_bases_upload): validated, Montgomery, [P | phi(P)] — nothing to convert.
static int msm_accumulate_phase(snarkv_ctx* ctx, const MsmWork& wk, const void* d_points, size_t n, int point_format, int check,
                                const void* d_points1 = nullptr, int add_existing = 0, uint8_t* conv_dst = nullptr, bool prepared = false) {
    const MsmPlan& pl = wk.pl;
    cudaStream_t st = ctx->stream;
    const size_t nbk = (size_t)pl.W * pl.NB;
    const int B = d_points1 ? 2 : 1;
    const size_t nv = plan_virtual_terms(pl, n);
    const uint8_t* pts[2] = {(const uint8_t*)d_points, (const uint8_t*)d_points1};
    if (!prepared && (point_format == SNARKV_CANONICAL || check || pl.glv)) {
        Stage sg(ctx, "msm_points_prepare");
        uint8_t* conv = nullptr;
        if (point_format == SNARKV_CANONICAL || pl.glv) {   // GLV gathers from one array [P | phi(P)] of nv points per base set
            conv = conv_dst ? conv_dst : (uint8_t*)ctx->wsget(WS_POINTS_MONT, (size_t)B * nv * 64);
            if (!conv) return SNARKV_ERR_CUDA;
        }
        const size_t want = (n + 255) / 256, cap = (size_t)ctx->sm_count * 8;
        for (int z = 0; z < B; ++z) {
            uint8_t* cz = conv ? conv + (size_t)z * nv * 64 : nullptr;
            k_points_prepare<<<(int)(want < cap ? want : cap), 256, 0, st>>>(pts[z], cz, n, point_format, check, (int)pl.glv, wk.status);
            SNARKV_LAUNCH_CHECK(ctx, "k_points_prepare");
            sg.launched();
            if (cz) pts[z] = cz;
        }
    }
    // accumulate mode.  Batched affine (6 instead of ~8.3 multiplications' worth of products per addition, but one inversion per
    // batch, a DRAM-bound forward pass, prefix traffic and whole 32-task units of work per warp) pays once the bucket lists are long.
    // Measured on B200 (tools/accumulate_probe.py) after the XYZZ addition got its dedicated squarings and the dual-product y3:
    // mean load 256 at 2^24 terms 33.5 vs 36.7 ms, 128 at 2^23 17.9 vs 18.3 ms; mean load 64 at 2^22 (c = 17) loses, 9.43 vs 9.13 ms,
    // and so do 128 at 2^21 (GLV) and every shorter list.  ba_min_load = 96 is the crossover for plans with >= 2^19 buckets.
    const int mode = ctx->accumulate_mode;
    const size_t mean_load = nv / pl.NB;
    bool affine = mode == 2 || mode == 3 || (mode == 0 && (mean_load >= 256 || (mean_load >= (size_t)ctx->ba_min_load && nbk >= ((size_t)1 << 19))));
    const size_t stride_a = (nv + pl.cap) / 2 + 2, stride_b = (stride_a + pl.cap) / 2 + 2;
    if (affine && mode == 0) {
        // the tree levels need scratch (48 B per term and window: 12 GB at 2^24 terms, 48 GB at 2^26); when that does not fit
        // next to the caller's data the automatic choice falls back to the XYZZ kernel, which needs none
        const size_t need = (size_t)B * pl.W * (stride_a + stride_b) * 64;
        const size_t held = ctx->ws_bytes[WS_BA_REGION_A] + ctx->ws_bytes[WS_BA_REGION_B];
        size_t free_b = 0, total_b = 0;
        if (need > held && (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess || need + need / 8 + ((size_t)2 << 30) > free_b + held)) affine = false;
    }
    // mode 4 (and the automatic choice once measured): chained batched-affine kernel (bucket_chain.cuh) — no scratch regions
    const bool chain = mode == 4 || mode == 5 || (mode == 0 && ctx->bc_auto && mean_load >= (size_t)ctx->bc_min_load);
    if (chain) {
        Stage sg(ctx, "msm_bucket_accumulate_chain");
        const int R = ctx->bc_r;
        auto kernel = R == 8 ? k_bucket_accumulate_chain<8> : R == 12 ? k_bucket_accumulate_chain<12> : k_bucket_accumulate_chain<16>;
        const size_t smem = (size_t)(SNARKV_BC_THREADS / 32) * R * 1024;
        if (ctx->bc_blocks_per_sm == 0 || ctx->bc_blocks_r != R) {
            SNARKV_CUDA_TRY(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            int per_sm = 0;
            SNARKV_CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, SNARKV_BC_THREADS, smem));
            ctx->bc_blocks_per_sm = per_sm > 0 ? per_sm : 1;
            ctx->bc_blocks_r = R;
        }
        const uint32_t units = ((pl.cap + 31) / 32) * pl.W * (uint32_t)B;   // 32 tasks each
        uint32_t blocks = (uint32_t)(ctx->sm_count * ctx->bc_blocks_per_sm);
        if (blocks > (units + 3) / 4) blocks = (units + 3) / 4;
        uint4* slab = (uint4*)ctx->wsget(WS_BC_SLAB, (size_t)ctx->sm_count * ctx->bc_blocks_per_sm * (SNARKV_BC_THREADS / 32) * R * 2048);
        uint32_t* ctr = (uint32_t*)ctx->wsget(WS_BA_COUNTER, 32);
        if (!slab || !ctr) return SNARKV_ERR_CUDA;
        SNARKV_CUDA_TRY(ctx, cudaMemsetAsync(ctr, 0, 32, st));
        kernel<<<blocks, SNARKV_BC_THREADS, smem, st>>>(pts[0], pts[1], wk.sorted, wk.offsets, wk.counts, wk.tasks, wk.window_tasks, wk.order, nv,
                                                        pl.NB, pl.T, pl.cap, pl.W, (uint32_t)B, wk.task_out, slab, ctr);
        SNARKV_LAUNCH_CHECK(ctx, "k_bucket_accumulate_chain");
        sg.launched();
        affine = false;
    }
    if ((!affine && !chain) || mode == 3 || mode == 5) {
        Stage sg(ctx, "msm_bucket_accumulate");
        uint8_t* dst = wk.task_out;
        if (mode == 3 || mode == 5) {
            dst = (uint8_t*)ctx->wsget(WS_TASK_OUT_CHECK, (size_t)B * pl.W * pl.cap * 128);
            if (!dst) return SNARKV_ERR_CUDA;
        }
        dim3 grid((pl.cap + 127) / 128, pl.W, B);
        k_bucket_accumulate<<<grid, 128, 0, st>>>(pts[0], pts[1], wk.sorted, wk.offsets, wk.counts, wk.tasks, wk.window_tasks, wk.order, nv,
                                                 pl.NB, pl.T, pl.cap, dst);
        SNARKV_LAUNCH_CHECK(ctx, "k_bucket_accumulate");
        sg.launched();
    }
    if (affine) {
        Stage sg(ctx, "msm_bucket_accumulate_affine");
        auto kernel = k_bucket_accumulate_affine;
        if (ctx->ba_blocks_per_sm == 0) {
            int per_sm = 0;
            SNARKV_CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, SNARKV_BA_THREADS, 0));
            ctx->ba_blocks_per_sm = per_sm > 0 ? per_sm : 1;
        }
        const uint32_t units = ((pl.cap + 31) / 32) * pl.W * (uint32_t)B;   // 32 tasks each; one warp takes up to Q at a time
        uint32_t blocks = (uint32_t)(ctx->sm_count * ctx->ba_blocks_per_sm);
        if (blocks > (units + 3) / 4) blocks = (units + 3) / 4;
        uint8_t* reg_a = (uint8_t*)ctx->wsget(WS_BA_REGION_A, (size_t)B * pl.W * stride_a * 64);
        uint8_t* reg_b = (uint8_t*)ctx->wsget(WS_BA_REGION_B, (size_t)B * pl.W * stride_b * 64);
        uint8_t* slab = (uint8_t*)ctx->wsget(WS_BA_PREFIX, (size_t)ctx->sm_count * ctx->ba_blocks_per_sm * ctx->ba_q * ctx->ba_k * SNARKV_BA_THREADS * 32);
        uint32_t* ctr = (uint32_t*)ctx->wsget(WS_BA_COUNTER, 32);
        if (!reg_a || !reg_b || !slab || !ctr) return SNARKV_ERR_CUDA;
        SNARKV_CUDA_TRY(ctx, cudaMemsetAsync(ctr, 0, 32, st));
        kernel<<<blocks, SNARKV_BA_THREADS, 0, st>>>(pts[0], pts[1], wk.sorted, wk.offsets, wk.counts, wk.tasks, wk.window_tasks,
                                                                        wk.order, nv, pl.NB, pl.T, pl.cap, pl.W, (uint32_t)B, wk.task_out, reg_a,
                                                                        reg_b, stride_a, stride_b, slab, ctr, (uint32_t)ctx->ba_k,
                                                                        (uint32_t)ctx->ba_pairs_min, (uint32_t)ctx->ba_q);
        SNARKV_LAUNCH_CHECK(ctx, "k_bucket_accumulate_affine");
        sg.launched();
    }
    if (mode == 3 || mode == 5) {   // debugging aid: both kernels ran; compare every task result as a group element (synchronous)
        Stage sg(ctx, "msm_bucket_accumulate_check");
        uint32_t* ctr = (uint32_t*)ctx->ws[WS_BA_COUNTER];
        dim3 grid((pl.cap + 127) / 128, pl.W, B);
        k_compare_task_results<<<grid, 128, 0, st>>>(wk.task_out, (const uint8_t*)ctx->ws[WS_TASK_OUT_CHECK], wk.window_tasks, wk.order, pl.cap,
                                                    pl.W, ctr + 1);
        SNARKV_LAUNCH_CHECK(ctx, "k_compare_task_results");
        sg.launched();
        uint32_t rec[4] = {};
        SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(rec, ctr + 1, 16, cudaMemcpyDeviceToHost, st));
        SNARKV_CUDA_TRY(ctx, cudaStreamSynchronize(st));
        if (rec[0] != 0) {
            char msg[160];
            snprintf(msg, sizeof msg, "batched-affine self-check: %u task result(s) differ from the XYZZ kernel; first: window %u, slot %u, set %u",
                     rec[0], rec[1], rec[2], rec[3]);
            return ctx->fail(SNARKV_ERR_CUDA, msg);
        }
    }
    {
        Stage sg(ctx, "msm_bucket_merge");
        const uint32_t total = (uint32_t)nbk;
        SNARKV_CUDA_TRY(ctx, cudaMemsetAsync(wk.big, 0, 4, st));
        k_bucket_merge<<<dim3((total + 127) / 128, B), 128, 0, st>>>(wk.task_out, wk.counts, wk.task_base, pl.NB, total, pl.T, pl.cap, wk.buckets,
                                                            wk.big, wk.big + 1, add_existing);
        SNARKV_LAUNCH_CHECK(ctx, "k_bucket_merge");
        sg.launched();
        k_bucket_merge_big<<<dim3(ctx->sm_count, B), 128, 0, st>>>(wk.task_out, wk.counts, wk.task_base, pl.NB, pl.T, pl.cap, wk.buckets,
                                                                   wk.big, wk.big + 1, pl.W, add_existing);
        SNARKV_LAUNCH_CHECK(ctx, "k_bucket_merge_big");
        sg.launched();
    }
    return SNARKV_OK;
}

// Point phase, second half: bucket reduction, window sums, Horner tail, normalisation.  Results land at d_out_affine + 64 z and
// d_out_jacobian + 96 z for base set z < B.
static int msm_tail_phase(snarkv_ctx* ctx, const MsmWork& wk, int B, int out_format, void* d_out_affine, void* d_out_jacobian) {
    const MsmPlan& pl = wk.pl;
    cudaStream_t st = ctx->stream;
    {
        Stage sg(ctx, "msm_bucket_reduce");
        dim3 grid((pl.J + 127) / 128, pl.W, B);
        k_bucket_reduce<<<grid, 128, 0, st>>>(wk.buckets, pl.NB, pl.seg, pl.J, wk.segpart);
        SNARKV_LAUNCH_CHECK(ctx, "k_bucket_reduce");
        sg.launched();
    }
    {
        Stage sg(ctx, "msm_window_sum");
        k_window_sum<<<dim3(pl.W, B), 128, 0, st>>>(wk.segpart, (pl.J + 127) / 128, wk.winsum);   // one partial per reduce block
        SNARKV_LAUNCH_CHECK(ctx, "k_window_sum");
        sg.launched();
    }
    {
        Stage sg(ctx, "msm_final");
        k_msm_final<<<B, 32, 0, st>>>(wk.winsum, pl.W, pl.c, out_format, d_out_affine, d_out_jacobian);
        SNARKV_LAUNCH_CHECK(ctx, "k_msm_final");
        sg.launched();
    }
    return SNARKV_OK;
}

static int msm_point_phase(snarkv_ctx* ctx, const MsmWork& wk, const void* d_points, size_t n, int point_format, int out_format,
                           int check, void* d_out_affine, void* d_out_jacobian, const void* d_points1 = nullptr) {
    int rc = msm_accumulate_phase(ctx, wk, d_points, n, point_format, check, d_points1);
    if (rc) return rc;
    return msm_tail_phase(ctx, wk, d_points1 ? 2 : 1, out_format, d_out_affine, d_out_jacobian);
}

// Two MSMs over the SAME scalars (KzgAs::verify: sum r^i lhs_i and sum r^i rhs_i, accumulation.rs:53-60): one digit/sort
// phase, every point-phase kernel launched once with a batch dimension.  d_out_affine receives 2 x 64 B.
int msm_run_device_pair(snarkv_ctx* ctx, const void* d_scalars, const void* d_points0, const void* d_points1, size_t n, int scalar_format,
                        int point_format, int out_format, int flags, void* d_out_affine, void* d_status) {
    const int check = (flags & SNARKV_CHECK_INPUTS) ? 1 : 0;
    MsmWork wk;
    int rc = msm_alloc(ctx, n, d_status, wk, 2);
    if (rc) return rc;
    rc = msm_sort_phase(ctx, wk, d_scalars, n, scalar_format, check);
    if (rc) return rc;
    return msm_point_phase(ctx, wk, d_points0, n, point_format, out_format, check, d_out_affine, nullptr, d_points1);
}

int msm_run_device(snarkv_ctx* ctx, const void* d_scalars, const void* d_points, size_t n, int scalar_format, int point_format,
                   int out_format, int flags, void* d_out_affine, void* d_out_jacobian, void* d_status) {
    const int check = (flags & SNARKV_CHECK_INPUTS) ? 1 : 0;
    MsmWork wk;
    int rc = msm_alloc(ctx, n, d_status, wk);
    if (rc) return rc;
    rc = msm_sort_phase(ctx, wk, d_scalars, n, scalar_format, check);
    if (rc) return rc;
    return msm_point_phase(ctx, wk, d_points, n, point_format, out_format, check, d_out_affine, d_out_jacobian);
}

// snarkv_g1_msm / snarkv_g1_msm_partial: host slices in.  Small inputs: scalars first, the (2x larger) point copy is issued
// after the sort kernels are queued so copy engine and SMs overlap.  Large inputs (>= 2^22 terms) are cut into 7 term-chunks of
// geometrically growing size (x1.6; measured sweep profiles/r01_host_chunks.txt): chunk k+1 is copied on the copy stream while chunk k is sorted and accumulated INTO THE SAME
// bucket array (plan fixed from the total n), so the first copy is short, no chunk pays its own reduce / Horner tail, and the
// copy engine stays ahead of the SMs (96 B/term at ~50 GB/s vs ~2.9 ns/term of compute).
// `d_resident` != nullptr: the bases already sit in HBM (prepared by msm_bases_prepare: Montgomery, validated, phi(P) appended);
// `points` is then ignored and only the scalars cross PCIe.
int msm_run_host(snarkv_ctx* ctx, const uint8_t* scalars, const uint8_t* points, size_t n, int format, int flags, uint8_t* out,
                 void* d_out_jacobian, const uint8_t* d_resident) {
    const int check = (flags & SNARKV_CHECK_INPUTS) ? 1 : 0;
    if (n == 0) return ctx->fail(SNARKV_ERR_EMPTY, "multi_scalar_multiplication on an empty slice");
    const bool resident = d_resident != nullptr;
    const int pfmt = resident ? SNARKV_MONTGOMERY : format;
    uint8_t* d_s = (uint8_t*)ctx->wsget(WS_IO_A, n * 32);
    uint8_t* d_p = resident ? const_cast<uint8_t*>(d_resident) : (uint8_t*)ctx->wsget(WS_IO_B, n * 64);
    uint8_t* d_o = (uint8_t*)ctx->wsget(WS_OUT, 1024);   // [affine 64 | pad | status K x 4 @512]
    if (!d_s || !d_p || !d_o) return SNARKV_ERR_CUDA;
    cudaStream_t st = ctx->stream;
    // Chunk pipeline: >= 2^22 terms in host_chunks (7) geometrically growing chunks; 2^20 .. 2^22 terms in 3.  On one GPU the short
    // pipeline is neutral at 2^21 terms (12.1 ms chunked vs 11.9 ms in one pass: per-chunk scans and merges eat the overlap), but when
    // the 8 GPUs of a box copy at the same time the host side of PCIe is the bottleneck and it pays: e2e of the 2^24-term MSM at
    // N = 8 (2^21 terms per rank) 1086 -> 1220 M/s (profiles/r02_bench_n8*.json).  SNARKV_HOST_CHUNK_MIN / _CHUNKS_SMALL override.
    const int K = n >= ((size_t)1 << ctx->host_chunk_min_log_n) ? (n >= ((size_t)1 << 22) ? ctx->host_chunks : ctx->host_chunks_small) : 1;
    int status[SNARKV_HOST_CHUNKS_MAX] = {};
    int* d_status = (int*)(d_o + 512);
    MsmWork wk;
    int rc = msm_alloc(ctx, n, d_status, wk, 1, K > 1 ? 2 : -1);   // plan from the TOTAL size; the chunk pipeline runs without GLV
    if (rc) return rc;
    // Every early return below first drains the copy stream: it may still be reading the caller's (possibly pinned) buffers.
    auto bail = [&](int code) {
        cudaStreamSynchronize(ctx->copy_stream);
        cudaStreamSynchronize(st);
        return code;
    };
    // the copy stream must not run ahead of work already queued on the compute stream that still reads the staging buffers
    SNARKV_CUDA_TRY(ctx, cudaEventRecord(ctx->copy_done[SNARKV_HOST_CHUNKS_MAX], st));
    SNARKV_CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_done[SNARKV_HOST_CHUNKS_MAX], 0));
    if (K == 1) {
        SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(d_s, scalars, n * 32, cudaMemcpyHostToDevice, st));
        rc = msm_sort_phase(ctx, wk, d_s, n, format, check);
        if (rc) return bail(rc);
        if (!resident) {
            cudaError_t ce = cudaMemcpyAsync(d_p, points, n * 64, cudaMemcpyHostToDevice, ctx->copy_stream);
            if (ce == cudaSuccess) ce = cudaEventRecord(ctx->copy_done[0], ctx->copy_stream);
            if (ce == cudaSuccess) ce = cudaStreamWaitEvent(st, ctx->copy_done[0], 0);
            if (ce != cudaSuccess) return bail(ctx->fail(SNARKV_ERR_CUDA, "host->device copy of the points", ce));
        }
        rc = msm_accumulate_phase(ctx, wk, d_p, n, pfmt, resident ? 0 : check, nullptr, 0, nullptr, resident);
        if (rc) return bail(rc);
    } else {
        size_t lo[SNARKV_HOST_CHUNKS_MAX + 1];
        const double ratio = ctx->host_chunk_ratio_pct / 100.0;
        double total_w = 0, wgt = 1;
        for (int k = 0; k < K; ++k, wgt *= ratio) total_w += wgt;
        double acc_w = 0;
        wgt = 1;
        lo[0] = 0;
        for (int k = 0; k < K; ++k, wgt *= ratio) {
            acc_w += wgt;
            lo[k + 1] = (k == K - 1) ? n : (((size_t)((double)n * (acc_w / total_w))) & ~(size_t)255);
        }
        uint8_t* conv = nullptr;
        if (format == SNARKV_CANONICAL && !resident) {
            conv = (uint8_t*)ctx->wsget(WS_POINTS_MONT, n * 64);
            if (!conv) return SNARKV_ERR_CUDA;
        }
        // Copies are issued one chunk AHEAD of the kernels that consume them (chunk k + 1 goes onto the copy stream right before
        // chunk k's kernels are launched), so that with pageable host memory — where cudaMemcpyAsync blocks the calling thread
        // while it stages — the SMs already have chunk k's work queued; real overlap needs PINNED host buffers.
        auto issue_copy = [&](int k) -> cudaError_t {
            const size_t len = lo[k + 1] - lo[k];
            cudaError_t ce = cudaMemcpyAsync(d_s + lo[k] * 32, scalars + lo[k] * 32, len * 32, cudaMemcpyHostToDevice, ctx->copy_stream);
            if (ce == cudaSuccess && !resident)
                ce = cudaMemcpyAsync(d_p + lo[k] * 64, points + lo[k] * 64, len * 64, cudaMemcpyHostToDevice, ctx->copy_stream);
            if (ce == cudaSuccess) ce = cudaEventRecord(ctx->copy_done[k], ctx->copy_stream);
            return ce;
        };
        cudaError_t ce = issue_copy(0);
        if (ce != cudaSuccess) return bail(ctx->fail(SNARKV_ERR_CUDA, "host->device copy (chunk 0)", ce));
        for (int k = 0; k < K; ++k) {
            const size_t len = lo[k + 1] - lo[k];
            if (k + 1 < K && (ce = issue_copy(k + 1)) != cudaSuccess) return bail(ctx->fail(SNARKV_ERR_CUDA, "host->device copy (chunk)", ce));
            if ((ce = cudaStreamWaitEvent(st, ctx->copy_done[k], 0)) != cudaSuccess) return bail(ctx->fail(SNARKV_ERR_CUDA, "cudaStreamWaitEvent", ce));
            MsmWork wc = wk;
            wc.status = d_status + k;
            rc = msm_sort_phase(ctx, wc, d_s + lo[k] * 32, len, format, check);
            if (rc) return bail(rc);
            rc = msm_accumulate_phase(ctx, wc, d_p + lo[k] * 64, len, pfmt, resident ? 0 : check, nullptr, k > 0, conv ? conv + lo[k] * 64 : nullptr,
                                      resident);
            if (rc) return bail(rc);
        }
    }
    rc = msm_tail_phase(ctx, wk, 1, format, out ? d_o : nullptr, d_out_jacobian);
    if (rc) return bail(rc);
    if (out) SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(out, d_o, 64, cudaMemcpyDeviceToHost, st));
    SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(status, d_status, 4 * K, cudaMemcpyDeviceToHost, st));
    SNARKV_CUDA_TRY(ctx, cudaStreamSynchronize(st));
    for (int k = 0; k < K; ++k)
        if (status[k] != 0)
            return ctx->fail(status[k], status[k] == SNARKV_ERR_BAD_SCALAR ? "scalar is not a canonical Fr" : "point is not a valid G1Affine");
    return SNARKV_OK;
}

// Resident base set: d_out (2 n x 64 B) <- [Montgomery P_i | phi(P_i)], validated when `check`; status word reports BAD_POINT.
int msm_bases_prepare(snarkv_ctx* ctx, const void* d_in, size_t n, int format, int check, void* d_out, void* d_status) {
    Stage sg(ctx, "msm_bases_prepare");
    SNARKV_CUDA_TRY(ctx, cudaMemsetAsync(d_status, 0, 4, ctx->stream));
    const size_t want = (n + 255) / 256, cap = (size_t)ctx->sm_count * 8;
    k_points_prepare<<<(int)(want < cap ? want : cap), 256, 0, ctx->stream>>>((const uint8_t*)d_in, (uint8_t*)d_out, n, format, check, 1,
                                                                             (int*)d_status);
    SNARKV_LAUNCH_CHECK(ctx, "k_points_prepare");
    sg.launched();
    return SNARKV_OK;
}

void msm_plan_query(snarkv_ctx* ctx, size_t n, uint32_t out[4]) {
    const MsmPlan pl = make_plan(n ? n : 1, ctx->window_bits, ctx->glv_mode);
    out[0] = pl.c; out[1] = pl.W; out[2] = pl.NB; out[3] = pl.T;
}

int msm_fold_partials_peer(snarkv_ctx* ctx, const void* const* d_partials, size_t k, int format, void* d_out_affine) {
    if (k == 0 || k > SNARKV_MAX_DEVICES) return ctx->fail(SNARKV_ERR_USAGE, "fold_partials_peer: bad k");
    PartialPtrs src = {};
    for (size_t i = 0; i < k; ++i) src.p[i] = (const uint8_t*)d_partials[i];
    Stage sg(ctx, "msm_fold_partials_peer");
    k_fold_partials_peer<<<1, 32, 0, ctx->stream>>>(src, (uint32_t)k, format, d_out_affine, nullptr);
    SNARKV_LAUNCH_CHECK(ctx, "k_fold_partials_peer");
    sg.launched();
    return SNARKV_OK;
}

int msm_fold_partials_device(snarkv_ctx* ctx, const void* d_partials, size_t k, int format, void* d_out_affine) {
    if (k == 0 || k > (1u << 20)) return ctx->fail(SNARKV_ERR_USAGE, "fold_partials: bad k");
    Stage sg(ctx, "msm_fold_partials");
    k_fold_partials<<<1, 32, 0, ctx->stream>>>((const uint8_t*)d_partials, (uint32_t)k, format, d_out_affine, nullptr);
    SNARKV_LAUNCH_CHECK(ctx, "k_fold_partials");
    sg.launched();
    return SNARKV_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// batch of small independent MSMs: the literal GPU analogue of loader/native.rs:61-71 — one thread per term computes
// base * scalar (double-and-add over the canonical bits, MSB first), one thread per MSM folds its terms and normalises.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_batch_term_mul(const uint8_t* __restrict__ scalars, const uint8_t* __restrict__ points,
                                                        size_t total, int format, int check, uint8_t* __restrict__ terms,
                                                        int* __restrict__ status) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    Fr s = fp_load<FR>(scalars + i * 32);
    G1Affine p = g1_affine_load(points, i);
    if (check && !fp_is_canonical(s)) atomicCAS(status, 0, SNARKV_ERR_BAD_SCALAR);
    if (check && (!fp_is_canonical(p.x) || !fp_is_canonical(p.y))) atomicCAS(status, 0, SNARKV_ERR_BAD_POINT);
    if (format == SNARKV_MONTGOMERY) s = fp_from_mont(s);
    else {
        p.x = fp_to_mont(p.x);
        p.y = fp_to_mont(p.y);
    }
    if (check && !g1_affine_is_on_curve(p)) atomicCAS(status, 0, SNARKV_ERR_BAD_POINT);
    G1Xyzz acc = xyzz_identity();
    if (!g1_affine_is_identity(p)) {
        for (int b = SNARKV_FIELD_BITS - 1; b >= 0; --b) {
            acc = xyzz_dbl(acc);
            if ((s.v[b >> 5] >> (b & 31)) & 1u) xyzz_madd(acc, p.x, p.y);
        }
    }
    xyzz_store(terms, i, acc);
}

__global__ void __launch_bounds__(128) k_batch_segment_sum(const uint8_t* __restrict__ terms, const uint64_t* __restrict__ offsets,
                                                           size_t m, int format, uint8_t* __restrict__ out) {
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    G1Xyzz acc = xyzz_identity();
    for (uint64_t i = offsets[j]; i < offsets[j + 1]; ++i) acc = xyzz_add(acc, xyzz_load(terms, i));
    store_affine_fmt(out + j * 64, xyzz_to_affine(acc), format);
}

int msm_batch_device(snarkv_ctx* ctx, const void* d_scalars, const void* d_points, const void* d_offsets, size_t m, size_t total,
                     int format, int flags, void* d_out_affine, void* d_status) {
    const int check = (flags & SNARKV_CHECK_INPUTS) ? 1 : 0;
    uint8_t* terms = (uint8_t*)ctx->wsget(WS_BATCH_TERMS, total * 128);
    if (!terms) return SNARKV_ERR_CUDA;
    SNARKV_CUDA_TRY(ctx, cudaMemsetAsync(d_status, 0, 4, ctx->stream));
    {
        Stage sg(ctx, "msm_batch_term_mul");
        k_batch_term_mul<<<(unsigned)((total + 127) / 128), 128, 0, ctx->stream>>>((const uint8_t*)d_scalars, (const uint8_t*)d_points,
                                                                                  total, format, check, terms, (int*)d_status);
        SNARKV_LAUNCH_CHECK(ctx, "k_batch_term_mul");
        sg.launched();
    }
    {
        Stage sg(ctx, "msm_batch_segment_sum");
        k_batch_segment_sum<<<(unsigned)((m + 127) / 128), 128, 0, ctx->stream>>>(terms, (const uint64_t*)d_offsets, m, format,
                                                                                 (uint8_t*)d_out_affine);
        SNARKV_LAUNCH_CHECK(ctx, "k_batch_segment_sum");
        sg.launched();
    }
    return SNARKV_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// r.powers(n)  (snark-verifier/src/loader.rs:71-78) on the device: out[i] = r^i in Montgomery form
// ---------------------------------------------------------------------------------------------------------------------
__global__ void k_fr_powers(const uint8_t* __restrict__ r_in, int format, size_t n, uint64_t first, uint8_t* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr r = fp_load<FR>(r_in);
    if (format == SNARKV_CANONICAL) r = fp_to_mont(r);
    Fr acc = fp_one<FR>();
    for (uint64_t e = first + i; e != 0; e >>= 1) {
        if (e & 1) acc = fp_mul(acc, r);
        r = fp_sqr(r);
    }
    fp_store<FR>(out + i * 32, acc);
}

// out[i] = r^(first + i): `first` > 0 is a shard of a longer power sequence (multi-device RLC batches, multi.cu)
int fr_powers_device(snarkv_ctx* ctx, const void* d_r, int format, size_t n, void* d_out_mont, uint64_t first) {
    Stage sg(ctx, "fr_powers");
    k_fr_powers<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>((const uint8_t*)d_r, format, n, first, (uint8_t*)d_out_mont);
    SNARKV_LAUNCH_CHECK(ctx, "k_fr_powers");
    sg.launched();
    return SNARKV_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// Fr vector helpers of the scalar-preparation step that precedes the MSM:
//   ScalarLoader::batch_invert (snark-verifier/src/loader.rs:255-262: zero stays zero) and util::arithmetic::
//   batch_invert_and_mul (util/arithmetic.rs:47-69: non-zero v -> coeff / v), element-wise products (RLC scalars rho^i * s_i).
// Batch inversion is Montgomery's trick per thread over a chunk of CHUNK values: prefix products of the non-zero values are parked
// in `scratch`, ONE inversion per chunk, then the backward sweep — 3 multiplications per value + 1/CHUNK of an inversion.
// ---------------------------------------------------------------------------------------------------------------------
#define SNARKV_INV_CHUNK 64
// `values` and `scratch` are written and read back by the same thread: plain (coherent) loads, no __restrict__ / __ldg on them.
__global__ void __launch_bounds__(128) k_fr_batch_invert(uint8_t* values, size_t n, int format, const uint8_t* __restrict__ coeff,
                                                         uint8_t* scratch) {
    const size_t chunk = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t lo = chunk * SNARKV_INV_CHUNK;
    if (lo >= n) return;
    const size_t hi = (lo + SNARKV_INV_CHUNK < n) ? lo + SNARKV_INV_CHUNK : n;
    Fr acc = fp_one<FR>();
    for (size_t i = lo; i < hi; ++i) {
        Fr v = fp_load_rw<FR>(values + i * 32);
        if (format == SNARKV_CANONICAL) v = fp_to_mont(v);
        fp_store<FR>(scratch + i * 32, acc);             // product of the non-zero values before i
        if (!fp_is_zero(v)) acc = fp_mul(acc, v);
    }
    Fr inv = fp_inv(acc);                                // acc != 0: product of non-zero field elements
    if (coeff) {
        Fr c = fp_load<FR>(coeff);
        if (format == SNARKV_CANONICAL) c = fp_to_mont(c);
        inv = fp_mul(inv, c);
    }
    for (size_t i = hi; i-- > lo;) {
        Fr v = fp_load_rw<FR>(values + i * 32);
        if (format == SNARKV_CANONICAL) v = fp_to_mont(v);
        if (fp_is_zero(v)) continue;                     // `unwrap_or_else(|| value.clone())`: zero is left as it is
        Fr out = fp_mul(inv, fp_load_rw<FR>(scratch + i * 32));
        inv = fp_mul(inv, v);
        if (format == SNARKV_CANONICAL) out = fp_from_mont(out);
        fp_store<FR>(values + i * 32, out);
    }
}

__global__ void __launch_bounds__(256) k_fr_mul_vec(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, size_t n, int format,
                                                    uint8_t* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr x = fp_load<FR>(a + i * 32), y = fp_load<FR>(b + i * 32);
    // canonical: (x R^-1-free) product needs one conversion: mont_mul(x, to_mont(y)) = x*y
    if (format == SNARKV_CANONICAL) y = fp_to_mont(y);
    fp_store<FR>(out + i * 32, fp_mul(x, y));            // Montgomery in -> Montgomery out; canonical in -> canonical out
}

__global__ void k_fr_from_mont(uint8_t* v, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) fp_store<FR>(v + i * 32, fp_from_mont(fp_load_rw<FR>(v + i * 32)));
}

int fr_batch_invert_device(snarkv_ctx* ctx, void* d_values, size_t n, int format, const void* d_coeff, void* d_scratch) {
    Stage sg(ctx, "fr_batch_invert");
    const size_t chunks = (n + SNARKV_INV_CHUNK - 1) / SNARKV_INV_CHUNK;
    k_fr_batch_invert<<<(unsigned)((chunks + 127) / 128), 128, 0, ctx->stream>>>((uint8_t*)d_values, n, format, (const uint8_t*)d_coeff,
                                                                              (uint8_t*)d_scratch);
    SNARKV_LAUNCH_CHECK(ctx, "k_fr_batch_invert");
    sg.launched();
    return SNARKV_OK;
}
int fr_mul_vec_device(snarkv_ctx* ctx, const void* d_a, const void* d_b, size_t n, int format, void* d_out) {
    Stage sg(ctx, "fr_mul_vec");
    k_fr_mul_vec<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>((const uint8_t*)d_a, (const uint8_t*)d_b, n, format, (uint8_t*)d_out);
    SNARKV_LAUNCH_CHECK(ctx, "k_fr_mul_vec");
    sg.launched();
    return SNARKV_OK;
}
int fr_from_mont_device(snarkv_ctx* ctx, void* d_v, size_t n) {
    k_fr_from_mont<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>((uint8_t*)d_v, n);
    SNARKV_LAUNCH_CHECK(ctx, "k_fr_from_mont");
    ctx->launches++;
    return SNARKV_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// RLC fusion of many MSMs into one (the batching of pcs/kzg/decider.rs:146-185 applied before the MSM instead of after it):
//   sum_j rho^j * MSM_j  =  one MSM over all terms with scalars  rho^j * s_ij .
// k_fr_scale_segments multiplies every scalar of segment j by rho^j (powers from k_fr_powers) and leaves Montgomery form.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_fr_scale_segments(const uint8_t* __restrict__ scalars, const uint64_t* __restrict__ offsets, size_t m,
                                                           size_t total, int format, int check, const uint8_t* __restrict__ powers_mont,
                                                           uint8_t* __restrict__ out_mont, int* __restrict__ status) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    size_t lo = 0, hi = m;   // largest j with offsets[j] <= i
    while (hi - lo > 1) {
        const size_t mid = (lo + hi) >> 1;
        if (offsets[mid] <= i) lo = mid; else hi = mid;
    }
    Fr s = fp_load<FR>(scalars + i * 32);
    if (check && !fp_is_canonical(s)) atomicCAS(status, 0, SNARKV_ERR_BAD_SCALAR);
    if (format == SNARKV_CANONICAL) s = fp_to_mont(s);
    fp_store<FR>(out_mont + i * 32, fp_mul(s, fp_load<FR>(powers_mont + lo * 32)));
}

int msm_batch_rlc_device(snarkv_ctx* ctx, const void* d_scalars, const void* d_points, const void* d_offsets, size_t m, size_t total,
                         const void* d_rho, int format, int flags, void* d_scaled /* total x 32 B scratch */, void* d_powers /* m x 32 B */,
                         void* d_out_affine, void* d_status, uint64_t first_power, void* d_out_jacobian) {
    const int check = (flags & SNARKV_CHECK_INPUTS) ? 1 : 0;
    int rc = fr_powers_device(ctx, d_rho, format, m, d_powers, first_power);
    if (rc) return rc;
    SNARKV_CUDA_TRY(ctx, cudaMemsetAsync(d_status, 0, 8, ctx->stream));
    {
        Stage sg(ctx, "fr_scale_segments");
        k_fr_scale_segments<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>((const uint8_t*)d_scalars, (const uint64_t*)d_offsets, m,
                                                                                      total, format, check, (const uint8_t*)d_powers,
                                                                                      (uint8_t*)d_scaled, (int*)d_status);
        SNARKV_LAUNCH_CHECK(ctx, "k_fr_scale_segments");
        sg.launched();
    }
    // the MSM keeps its own status word right after ours so that a scalar error found above is not overwritten
    return msm_run_device(ctx, d_scaled, d_points, total, SNARKV_MONTGOMERY, format, format, flags, d_out_affine, d_out_jacobian,
                          (uint8_t*)d_status + 4);
}

// ---------------------------------------------------------------------------------------------------------------------
// element-wise field operations (test support for the parity suite: pins fp.cuh against the golden field vectors)
// ---------------------------------------------------------------------------------------------------------------------
template <Field F>
__global__ void k_field_op(int op, const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, size_t n, uint8_t* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fp<F> x = fp_to_mont(fp_load<F>(a + i * 32));
    Fp<F> y = fp_to_mont(fp_load<F>(b + i * 32));
    Fp<F> r;
    switch (op) {
        case 0: r = fp_mul(x, y); break;
        case 1: r = fp_add(x, y); break;
        case 2: r = fp_sub(x, y); break;
        case 3: r = fp_inv(x); break;
        default: r = fp_sqr(x); break;
    }
    fp_store<F>(out + i * 32, fp_from_mont(r));
}

int field_op_device(snarkv_ctx* ctx, int field, int op, const void* d_a, const void* d_b, size_t n, void* d_out) {
    const unsigned blocks = (unsigned)((n + 127) / 128);
    if (field == 0) k_field_op<FQ><<<blocks, 128, 0, ctx->stream>>>(op, (const uint8_t*)d_a, (const uint8_t*)d_b, n, (uint8_t*)d_out);
    else k_field_op<FR><<<blocks, 128, 0, ctx->stream>>>(op, (const uint8_t*)d_a, (const uint8_t*)d_b, n, (uint8_t*)d_out);
    SNARKV_LAUNCH_CHECK(ctx, "k_field_op");
    ctx->launches++;
    return SNARKV_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// synthetic workload generators (definition in include/snarkv_cuda.h; the test suite restates it independently)
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void k_synth_scalars(uint64_t seed, uint64_t start, size_t n, int format, uint8_t* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr s;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        uint64_t l = splitmix64(seed * 0x100000001B3ull + (start + i) * 4 + k);
        if (k == 3) l &= (1ull << 62) - 1;
        s.v[2 * k] = (uint32_t)l;
        s.v[2 * k + 1] = (uint32_t)(l >> 32);
    }
    if (!fp_is_canonical(s)) {  // one conditional subtraction of r (value < 2^254 < 2r)
        uint32_t borrow = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            uint64_t d = (uint64_t)s.v[k] - fp_mod_limb<FR>(k) - borrow;
            s.v[k] = (uint32_t)d;
            borrow = (uint32_t)(d >> 63);
        }
    }
    if (format == SNARKV_MONTGOMERY) s = fp_to_mont(s);
    fp_store<FR>(out + i * 32, s);
}

__global__ void __launch_bounds__(128) k_synth_points(uint64_t seed, uint64_t start, size_t n, int format, uint8_t* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t t = splitmix64(seed * 0x100000001B3ull + 0x5151515151515151ull + start + i) | 1ull;
    Fq gx = fp_one<FQ>();
    Fq gy = fp_dbl(gx);
#ifdef SNARKV_CURVE_PALLAS
    gx = fp_neg(gx);      // Pallas generator (-1, 2); BN254 G1: (1, 2)
#endif
    G1Xyzz acc = xyzz_identity();
    for (int b = 63 - __clzll((long long)t); b >= 0; --b) {
        acc = xyzz_dbl(acc);
        if ((t >> b) & 1ull) xyzz_madd(acc, gx, gy);
    }
    store_affine_fmt(out + i * 64, xyzz_to_affine(acc), format);
}

int synth_scalars_device(snarkv_ctx* ctx, uint64_t seed, uint64_t start, size_t n, int format, void* d_out) {
    if (n == 0) return SNARKV_OK;
    k_synth_scalars<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(seed, start, n, format, (uint8_t*)d_out);
    SNARKV_LAUNCH_CHECK(ctx, "k_synth_scalars");
    ctx->launches++;
    return SNARKV_OK;
}
int synth_points_device(snarkv_ctx* ctx, uint64_t seed, uint64_t start, size_t n, int format, void* d_out) {
    if (n == 0) return SNARKV_OK;
    k_synth_points<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(seed, start, n, format, (uint8_t*)d_out);
    SNARKV_LAUNCH_CHECK(ctx, "k_synth_points");
    ctx->launches++;
    return SNARKV_OK;
}

}  // namespace snarkv
