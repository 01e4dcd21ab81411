// sort2.cuh — two-level bucket sort for the MSM sort phase (sort mode 1; EXPERIMENTAL, off by default).
//
// The single-level path (msm.cu K1 + K3) pays two rounds of L2 atomics per (term, window) — a histogram increment and a cursor
// bump on one of 2^(c-1) counters — and one random 4-byte store; at 2^24 terms that is 1.4 + 3.2 ms, 12 % of the step, at
// 17 % / 10 % of the HBM roofline on its algorithmic bytes.  Here the bucket id is split into a partition (high bits) and a low
// key of LB <= 8 bits:
//   level 1  k_part_count    per tile of 4096 digits: partition histogram in shared memory      -> tilecnt[w][p][tile]
//            k_part_scan     exclusive scan over the tiles of every (window, partition)           (in place) + partition totals
//            k_part_base     exclusive scan over the partitions of a window                      -> partbase[w][0..P]
//            k_part_scatter  per tile: shared-memory cursors, (term reference, low key) written in partition order
//   level 2  k_bin_count     one block per (window, partition): shared-memory histogram of the low keys = the BUCKET COUNTS
//            (the existing K2 scans turn the counts into bucket offsets and task descriptors, unchanged)
//            k_bin_place     one block per (window, partition): shared-memory cursors, references placed into the bucket runs
// All atomics are on shared memory; the global stores of level 1 are runs of ~16 consecutive elements, those of level 2 stay
// inside the partition's 256 KB slice.  Skewed inputs (every term in one bucket) stay correct: one block then walks a whole
// window with warp-aggregated atomics (__match_any_sync).  The index arithmetic is restated in numpy and checked against a plain
// sort in tests/test_two_level_sort_model.py.
#pragma once
#include <cstdint>

namespace snarkv {

#define SNARKV_S2_TILE 4096
#define SNARKV_S2_MAXP 256   // partitions per window (c <= 17: 16 bucket bits = 8 high + 8 low)

__global__ void __launch_bounds__(256) k_part_count(const uint32_t* __restrict__ digits, size_t nv, uint32_t lb, uint32_t P, uint32_t T,
                                                    uint32_t* __restrict__ tilecnt) {
    __shared__ uint32_t hist[SNARKV_S2_MAXP];
    const uint32_t tile = blockIdx.x, w = blockIdx.y, t = threadIdx.x;
    for (uint32_t p = t; p < P; p += blockDim.x) hist[p] = 0;
    __syncthreads();
    const uint32_t* dg = digits + (size_t)w * nv;
    const size_t base = (size_t)tile * SNARKV_S2_TILE;
#pragma unroll 4
    for (uint32_t j = 0; j < SNARKV_S2_TILE / 256; ++j) {
        const size_t i = base + (size_t)j * 256 + t;
        if (i < nv) {
            const uint32_t d = dg[i] & 0x7fffffffu;
            if (d) atomicAdd(&hist[(d - 1u) >> lb], 1u);
        }
    }
    __syncthreads();
    for (uint32_t p = t; p < P; p += blockDim.x) tilecnt[((size_t)w * P + p) * T + tile] = hist[p];
}

// in-place exclusive scan of the T tile counts of one (window, partition); the total goes to parttot
__global__ void __launch_bounds__(1024) k_part_scan(uint32_t* __restrict__ tilecnt, uint32_t P, uint32_t T, uint32_t* __restrict__ parttot) {
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t carry_sm;
    const uint32_t p = blockIdx.x, w = blockIdx.y, t = threadIdx.x, lane = t & 31u, wid = t >> 5;
    uint32_t* row = tilecnt + ((size_t)w * P + p) * T;
    if (t == 0) carry_sm = 0;
    __syncthreads();
    for (uint32_t t0 = 0; t0 < T; t0 += blockDim.x) {
        const uint32_t k = t0 + t;
        const uint32_t v = k < T ? row[k] : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= (uint32_t)o) incl += u;
        }
        if (lane == 31) warp_tot[wid] = incl;
        const uint32_t carry = carry_sm;   // read before the barrier that lets warp 0 overwrite it
        __syncthreads();
        if (wid == 0) {
            const uint32_t x = lane < (blockDim.x >> 5) ? warp_tot[lane] : 0u;
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
        if (k < T) row[k] = carry + warp_tot[wid] + incl - v;
        __syncthreads();
    }
    if (t == 0) parttot[(size_t)w * P + p] = carry_sm;
}

// partbase[w][p] = first slot of partition p in window w's partition-ordered stream; partbase[w][P] = number of non-zero digits
__global__ void __launch_bounds__(SNARKV_S2_MAXP) k_part_base(const uint32_t* __restrict__ parttot, uint32_t P, uint32_t* __restrict__ partbase) {
    __shared__ uint32_t warp_tot[SNARKV_S2_MAXP / 32];
    const uint32_t w = blockIdx.x, t = threadIdx.x, lane = t & 31u, wid = t >> 5;
    const uint32_t v = t < P ? parttot[(size_t)w * P + t] : 0u;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (uint32_t)o) incl += u;
    }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    uint32_t before = 0;
    for (uint32_t k = 0; k < wid; ++k) before += warp_tot[k];
    if (t < P) partbase[(size_t)w * (P + 1) + t] = before + incl - v;
    if (t == blockDim.x - 1) partbase[(size_t)w * (P + 1) + P] = before + incl;
}

__global__ void __launch_bounds__(256) k_part_scatter(const uint32_t* __restrict__ digits, size_t nv, uint32_t lb, uint32_t P, uint32_t T,
                                                      const uint32_t* __restrict__ tilebase, const uint32_t* __restrict__ partbase,
                                                      uint32_t* __restrict__ rec, uint8_t* __restrict__ key) {
    __shared__ uint32_t cur[SNARKV_S2_MAXP];
    const uint32_t tile = blockIdx.x, w = blockIdx.y, t = threadIdx.x;
    for (uint32_t p = t; p < P; p += blockDim.x) cur[p] = partbase[(size_t)w * (P + 1) + p] + tilebase[((size_t)w * P + p) * T + tile];
    __syncthreads();
    const uint32_t* dg = digits + (size_t)w * nv;
    uint32_t* ro = rec + (size_t)w * nv;
    uint8_t* ko = key + (size_t)w * nv;
    const size_t base = (size_t)tile * SNARKV_S2_TILE;
    const uint32_t lowmask = (1u << lb) - 1u;
#pragma unroll 4
    for (uint32_t j = 0; j < SNARKV_S2_TILE / 256; ++j) {
        const size_t i = base + (size_t)j * 256 + t;
        if (i < nv) {
            const uint32_t e = dg[i];
            const uint32_t d = e & 0x7fffffffu;
            if (d) {
                const uint32_t pos = atomicAdd(&cur[(d - 1u) >> lb], 1u);
                ro[pos] = (uint32_t)i | (e & 0x80000000u);
                ko[pos] = (uint8_t)((d - 1u) & lowmask);
            }
        }
    }
}

// bucket counts of one (window, partition): counts[w][p << lb | k] for the 2^lb low keys k
__global__ void __launch_bounds__(1024) k_bin_count(const uint8_t* __restrict__ key, size_t nv, uint32_t lb, uint32_t P, uint32_t NB,
                                                    const uint32_t* __restrict__ partbase, uint32_t* __restrict__ counts) {
    __shared__ uint32_t hist[256];
    const uint32_t p = blockIdx.x, w = blockIdx.y, t = threadIdx.x, lane = t & 31u;
    if (t < 256) hist[t] = 0;
    __syncthreads();
    const uint32_t lo = partbase[(size_t)w * (P + 1) + p], hi = partbase[(size_t)w * (P + 1) + p + 1];
    const uint8_t* kk = key + (size_t)w * nv;
    for (uint32_t j = lo + t; j < hi; j += blockDim.x) {
        const uint32_t k = kk[j];
        const uint32_t peers = __match_any_sync(__activemask(), k);   // one atomic per distinct key and warp (skew-proof)
        if (lane == (uint32_t)(__ffs(peers) - 1)) atomicAdd(&hist[k], (uint32_t)__popc(peers));
    }
    __syncthreads();
    if (t < (1u << lb)) counts[(size_t)w * NB + ((size_t)p << lb) + t] = hist[t];
}

// references of one (window, partition) placed into their bucket runs (offsets from the K2 scan of the counts above)
__global__ void __launch_bounds__(1024) k_bin_place(const uint8_t* __restrict__ key, const uint32_t* __restrict__ rec, size_t nv, uint32_t lb,
                                                    uint32_t P, uint32_t NB, const uint32_t* __restrict__ partbase,
                                                    const uint32_t* __restrict__ offsets, uint32_t* __restrict__ sorted) {
    __shared__ uint32_t cur[256];
    const uint32_t p = blockIdx.x, w = blockIdx.y, t = threadIdx.x, lane = t & 31u;
    if (t < (1u << lb)) cur[t] = offsets[(size_t)w * NB + ((size_t)p << lb) + t];
    __syncthreads();
    const uint32_t lo = partbase[(size_t)w * (P + 1) + p], hi = partbase[(size_t)w * (P + 1) + p + 1];
    const uint8_t* kk = key + (size_t)w * nv;
    const uint32_t* rr = rec + (size_t)w * nv;
    uint32_t* out = sorted + (size_t)w * nv;
    for (uint32_t j = lo + t; j < hi; j += blockDim.x) {
        const uint32_t k = kk[j];
        const uint32_t r = rr[j];
        const uint32_t peers = __match_any_sync(__activemask(), k);
        const uint32_t leader = (uint32_t)(__ffs(peers) - 1);
        uint32_t first = 0;
        if (lane == leader) first = atomicAdd(&cur[k], (uint32_t)__popc(peers));
        first = __shfl_sync(peers, first, leader);
        out[first + __popc(peers & ((1u << lane) - 1u))] = r;
    }
}

}  // namespace snarkv
