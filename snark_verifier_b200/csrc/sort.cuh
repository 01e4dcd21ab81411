// sort.cuh — the counting sort of the MSM's signed digits into bucket runs, large-input path.
//
// What it produces is what the accumulation kernels consume (msm.cu): for every window w the term references
// `index | sign << 31` grouped by bucket (`sorted`), plus the bucket sizes (`counts`).  The reference's Pippenger does this
// implicitly, one window at a time (`buckets[scalar - 1].add_assign(base)`, snark-verifier/src/util/msm.rs:291-296).
//
// The small-input path (k_digits histogram + k_scatter, msm.cu) pays one global atomic per (term, window) twice; beyond ~2^18
// terms those 2 n W L2 atomics and the n W random 4-byte stores are the whole cost (r01: 4.55 ms of a 39.9 ms step at 2^24, 17 % /
// 10 % of the HBM roofline).  This path is a two-level partition with NO global atomic per element:
//
//   K1 k_digits<PART>   scalars -> digits[w][i] (coalesced), histogram of the COARSE partition p = (d - 1) >> lo_bits kept in shared
//                       memory per block and flushed once per block                                           (HBM: 32 + 4 W B / term)
//   K2 k_part_scan      exclusive scan of the W x P partition sizes                                           (tiny)
//   K3 k_partition      work item = (window, tile of 4096 digits): rank inside the tile by shared-memory atomics, ONE global atomic
//                       per non-empty (tile, partition) reserves the run, records staged in shared memory in partition order and
//                       written as runs: rec_idx (4 B) / rec_lo (2 B) per record                              (HBM: 4 + 6 B / digit)
//   K4 k_sort_buckets   one block per (window, partition): histogram of the low digit bits in shared memory -> bucket sizes; the
//                       partition's output (<= 48 K references) is assembled in shared memory and written fully coalesced
//                                                                                                             (HBM: 6 + 4 B / digit)
// Skewed inputs stay correct: a partition larger than the shared-memory buffer is emitted in several bucket ranges, a single
// bucket larger than the buffer is written straight to global memory.
#pragma once
#include <cstdint>

namespace snarkv {

#define SNARKV_SORT_TILE 4096          // digits per k_partition work item
#define SNARKV_SORT_THREADS 256
#define SNARKV_SORT_PER_THREAD (SNARKV_SORT_TILE / SNARKV_SORT_THREADS)
#define SNARKV_SORT_CAP (44 * 1024)    // references one k_sort_buckets block can stage in shared memory (176 KB)
#define SNARKV_SORT_BUCKET_THREADS 1024

// K2: one block; exclusive scan of every window's P partition sizes -> part_off[w][p]; zeroes the run cursors
__global__ void __launch_bounds__(1024) k_part_scan(const uint32_t* __restrict__ part_count, uint32_t* __restrict__ part_off,
                                                    uint32_t* __restrict__ part_cursor, uint32_t W, uint32_t P) {
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t carry_sm;
    const uint32_t t = threadIdx.x, lane = t & 31, wid = t >> 5;
    for (uint32_t w = blockIdx.x; w < W; w += gridDim.x) {
        if (t == 0) carry_sm = 0;
        __syncthreads();
        for (uint32_t tile = 0; tile < P; tile += blockDim.x) {
            const uint32_t k = tile + t;
            const uint32_t v = k < P ? part_count[w * P + k] : 0u;
            uint32_t incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= (uint32_t)o) incl += u;
            }
            if (lane == 31) warp_tot[wid] = incl;
            const uint32_t carry = carry_sm;
            __syncthreads();
            if (wid == 0) {
                const uint32_t x = warp_tot[lane];
                uint32_t ix = x;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t u = __shfl_up_sync(0xffffffffu, ix, o);
                    if (lane >= (uint32_t)o) ix += u;
                }
                warp_tot[lane] = ix - x;
                if (lane == 31) carry_sm = carry + ix;
            }
            __syncthreads();
            if (k < P) {
                part_off[w * P + k] = carry + warp_tot[wid] + incl - v;
                part_cursor[w * P + k] = 0;
            }
            __syncthreads();
        }
    }
}

// block-wide exclusive scan of `len` (<= 8192) u32 values in shared memory, in place; returns the total.  All threads call it.
__device__ __forceinline__ uint32_t block_exclusive_scan_smem(uint32_t* a, uint32_t len, uint32_t* warp_tot /* >= 33 */) {
    const uint32_t t = threadIdx.x, nt = blockDim.x, lane = t & 31, wid = t >> 5;
    const uint32_t per = (len + nt - 1) / nt;          // consecutive entries per thread
    const uint32_t lo = t * per, hi = min(lo + per, len);
    uint32_t sum = 0;
    for (uint32_t i = lo; i < hi; ++i) sum += a[i];
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (uint32_t)o) incl += u;
    }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        const uint32_t x = lane < (nt >> 5) ? warp_tot[lane] : 0u;
        uint32_t ix = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, ix, o);
            if (lane >= (uint32_t)o) ix += u;
        }
        warp_tot[lane] = ix - x;
        if (lane == 31) warp_tot[32] = ix;
    }
    __syncthreads();
    uint32_t run = warp_tot[wid] + incl - sum;
    for (uint32_t i = lo; i < hi; ++i) {
        const uint32_t v = a[i];
        a[i] = run;
        run += v;
    }
    const uint32_t total = warp_tot[32];
    __syncthreads();
    return total;
}

// K3: coarse partition.  Dynamic shared memory: hist[P] | loff[P] | gbase[P] | stage_idx[TILE] | stage_lo[TILE] (u16) | stage_p[TILE] (u16)
template <int THREADS>   // tile = THREADS x SNARKV_SORT_PER_THREAD digits
__global__ void __launch_bounds__(THREADS) k_partition(const uint32_t* __restrict__ digits, size_t nv, uint32_t W, uint32_t P,
                                                                   uint32_t lo_bits, const uint32_t* __restrict__ part_off,
                                                                   uint32_t* __restrict__ part_cursor, uint32_t* __restrict__ rec_idx,
                                                                   uint16_t* __restrict__ rec_lo) {
    extern __shared__ __align__(16) uint8_t sort_smem[];
    __shared__ uint32_t warp_tot[33];
    uint32_t* hist = reinterpret_cast<uint32_t*>(sort_smem);
    uint32_t* loff = hist + P;
    uint32_t* gbase = loff + P;
    uint32_t* stage_idx = gbase + P;
    uint16_t* stage_lo = reinterpret_cast<uint16_t*>(stage_idx + (THREADS * SNARKV_SORT_PER_THREAD));
    uint16_t* stage_p = stage_lo + (THREADS * SNARKV_SORT_PER_THREAD);
    const uint32_t t = threadIdx.x;
    const size_t ntiles = (nv + (THREADS * SNARKV_SORT_PER_THREAD) - 1) / (THREADS * SNARKV_SORT_PER_THREAD);
    const size_t items = ntiles * W;
    const uint32_t lo_mask = (1u << lo_bits) - 1u;
    for (size_t item = blockIdx.x; item < items; item += gridDim.x) {
        const uint32_t w = (uint32_t)(item / ntiles);          // window-major: neighbouring blocks share a window's cursors
        const size_t first = (item - (size_t)w * ntiles) * (THREADS * SNARKV_SORT_PER_THREAD);
        const uint32_t* dg = digits + (size_t)w * nv;
        for (uint32_t k = t; k < P; k += THREADS) hist[k] = 0;
        __syncthreads();
        uint32_t e[SNARKV_SORT_PER_THREAD], rank[SNARKV_SORT_PER_THREAD];
#pragma unroll
        for (int k = 0; k < SNARKV_SORT_PER_THREAD; ++k) {
            const size_t i = first + (size_t)k * THREADS + t;
            e[k] = i < nv ? dg[i] : 0u;
        }
#pragma unroll
        for (int k = 0; k < SNARKV_SORT_PER_THREAD; ++k) {
            const uint32_t d = e[k] & 0x7fffffffu;
            rank[k] = d ? atomicAdd(&hist[(d - 1u) >> lo_bits], 1u) : 0u;
        }
        __syncthreads();
        // counts -> exclusive offsets inside the tile (loff) and the global run reserved for this tile (gbase)
        for (uint32_t k = t; k < P; k += THREADS) {
            const uint32_t c = hist[k];
            loff[k] = c;
            gbase[k] = c ? part_off[w * P + k] + atomicAdd(&part_cursor[w * P + k], c) : 0u;
        }
        __syncthreads();
        const uint32_t total = block_exclusive_scan_smem(loff, P, warp_tot);
#pragma unroll
        for (int k = 0; k < SNARKV_SORT_PER_THREAD; ++k) {
            const uint32_t d = e[k] & 0x7fffffffu;
            if (d == 0) continue;
            const uint32_t b = d - 1u, p = b >> lo_bits;
            const uint32_t slot = loff[p] + rank[k];
            const size_t i = first + (size_t)k * THREADS + t;
            stage_idx[slot] = (uint32_t)i | (e[k] & 0x80000000u);
            stage_lo[slot] = (uint16_t)(b & lo_mask);
            stage_p[slot] = (uint16_t)p;
        }
        __syncthreads();
        uint32_t* out_idx = rec_idx + (size_t)w * nv;
        uint16_t* out_lo = rec_lo + (size_t)w * nv;
        for (uint32_t slot = t; slot < total; slot += THREADS) {
            const uint32_t p = stage_p[slot];
            const uint32_t pos = gbase[p] + (slot - loff[p]);
            out_idx[pos] = stage_idx[slot];
            out_lo[pos] = stage_lo[slot];
        }
        __syncthreads();
    }
}

// K4: one block per (window, partition).  Dynamic shared memory: hist[LB] | start[LB + 1] | out[CAP]
// The records are read as aligned 16-byte vectors (8 low-digit values / 4 references per load), four vectors in flight per thread
// before the first shared-memory atomic: the first version issued one 2-byte load per atomic and was bound by global-load latency
// (3.4 ms at 2^24 terms).  `rec_idx` / `rec_lo` must be 32-byte aligned allocations with 32 bytes of slack at the end.
struct SortRecVec {
    uint32_t lo[4];    // 8 x u16
    uint32_t idx[8];
};
template <bool WANT_IDX>
__device__ __forceinline__ void sort_load_vec(const uint4* __restrict__ lov, const uint4* __restrict__ idxv, uint32_t v, SortRecVec& r) {
    const uint4 q = __ldg(lov + v);
    r.lo[0] = q.x; r.lo[1] = q.y; r.lo[2] = q.z; r.lo[3] = q.w;
    if (WANT_IDX) {
        const uint4 a = __ldg(idxv + 2 * (size_t)v), b = __ldg(idxv + 2 * (size_t)v + 1);
        r.idx[0] = a.x; r.idx[1] = a.y; r.idx[2] = a.z; r.idx[3] = a.w;
        r.idx[4] = b.x; r.idx[5] = b.y; r.idx[6] = b.z; r.idx[7] = b.w;
    }
}
#define SNARKV_SORT_MLP 4
__global__ void __launch_bounds__(SNARKV_SORT_BUCKET_THREADS) k_sort_buckets(const uint32_t* __restrict__ rec_idx, const uint16_t* __restrict__ rec_lo,
                                                                            size_t nv, uint32_t P, uint32_t lo_bits, uint32_t NB,
                                                                            const uint32_t* __restrict__ part_off,
                                                                            const uint32_t* __restrict__ part_count, uint32_t* __restrict__ counts,
                                                                            uint32_t* __restrict__ sorted) {
    extern __shared__ __align__(16) uint8_t sort_smem[];
    __shared__ uint32_t warp_tot[33];
    __shared__ uint32_t range_hi;
    const uint32_t LB = 1u << lo_bits;
    uint32_t* cur = reinterpret_cast<uint32_t*>(sort_smem);   // histogram, then per-bucket write cursors
    uint32_t* start = cur + LB;                               // exclusive bucket offsets inside the partition, start[LB] = S
    uint32_t* out = start + LB + 1;
    const uint32_t t = threadIdx.x, nt = blockDim.x;
    const uint32_t w = blockIdx.y, p = blockIdx.x;
    const uint32_t S = part_count[w * P + p];
    const uint32_t base = part_off[w * P + p];
    const size_t first = (size_t)w * nv + base;               // element index of the partition's first record
    const uint32_t mis = (uint32_t)(first & 7u);
    const uint4* lov = reinterpret_cast<const uint4*>(rec_lo + (first - mis));
    const uint4* idxv = reinterpret_cast<const uint4*>(rec_idx + (first - mis));
    const uint32_t nvec = (S + mis + 7u) >> 3;
    uint32_t* dst = sorted + first;
    for (uint32_t k = t; k < LB; k += nt) cur[k] = 0;
    __syncthreads();
    for (uint32_t v0 = t; v0 < nvec; v0 += nt * SNARKV_SORT_MLP) {
        SortRecVec r[SNARKV_SORT_MLP];
#pragma unroll
        for (int u = 0; u < SNARKV_SORT_MLP; ++u)
            if (v0 + u * nt < nvec) sort_load_vec<false>(lov, idxv, v0 + u * nt, r[u]);
#pragma unroll
        for (int u = 0; u < SNARKV_SORT_MLP; ++u) {
            const uint32_t v = v0 + u * nt;
            if (v >= nvec) break;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint32_t j = v * 8u + k - mis;          // wraps for the (< 8) elements before the partition: j >= S
                if (j < S) atomicAdd(&cur[(r[u].lo[k >> 1] >> (16 * (k & 1))) & 0xffffu], 1u);
            }
        }
    }
    __syncthreads();
    for (uint32_t k = t; k < LB; k += nt) {
        const uint32_t c = cur[k];
        counts[(size_t)w * NB + (size_t)p * LB + k] = c;      // bucket value d = p * LB + k + 1
        start[k] = c;
    }
    __syncthreads();
    block_exclusive_scan_smem(start, LB, warp_tot);
    if (t == 0) start[LB] = S;
    __syncthreads();
    // emit bucket ranges [l0, l1) whose references fit the shared buffer; a single oversized bucket goes straight to global memory
    uint32_t l0 = 0;
    while (l0 < LB) {
        if (t == 0) {
            uint32_t a = l0 + 1, b = LB;                       // largest l1 in (l0, LB] with start[l1] - start[l0] <= CAP, at least l0 + 1
            const uint32_t s0 = start[l0];
            while (a < b) {
                const uint32_t mid = (a + b + 1) >> 1;
                if (start[mid] - s0 <= SNARKV_SORT_CAP) a = mid; else b = mid - 1;
            }
            range_hi = a;
        }
        __syncthreads();
        const uint32_t l1 = range_hi;
        const uint32_t s0 = start[l0], cnt = start[l1] - s0;
        const bool staged = cnt <= SNARKV_SORT_CAP;
        for (uint32_t k = l0 + t; k < l1; k += nt) cur[k] = start[k];
        __syncthreads();
        if (cnt != 0) {
            for (uint32_t v0 = t; v0 < nvec; v0 += nt * SNARKV_SORT_MLP) {
                SortRecVec r[SNARKV_SORT_MLP];
#pragma unroll
                for (int u = 0; u < SNARKV_SORT_MLP; ++u)
                    if (v0 + u * nt < nvec) sort_load_vec<true>(lov, idxv, v0 + u * nt, r[u]);
#pragma unroll
                for (int u = 0; u < SNARKV_SORT_MLP; ++u) {
                    const uint32_t v = v0 + u * nt;
                    if (v >= nvec) break;
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const uint32_t j = v * 8u + k - mis;
                        const uint32_t l = (r[u].lo[k >> 1] >> (16 * (k & 1))) & 0xffffu;
                        if (j >= S || l < l0 || l >= l1) continue;
                        const uint32_t pos = atomicAdd(&cur[l], 1u);
                        if (staged) out[pos - s0] = r[u].idx[k];
                        else dst[pos] = r[u].idx[k];
                    }
                }
            }
            __syncthreads();
            if (staged)
                for (uint32_t j = t; j < cnt; j += nt) dst[s0 + j] = out[j];
        }
        __syncthreads();
        l0 = l1;
    }
}

}  // namespace snarkv
