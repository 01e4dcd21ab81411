// bucket_affine.cuh — batched-affine bucket accumulation: the same `buckets[scalar - 1].add_assign(base)` of
// snark-verifier/src/util/msm.rs:291-296 as k_bucket_accumulate (msm.cu), with 6 instead of 10 Montgomery multiplications
// per point addition.
//
// A task is a run of <= T sorted point references of one bucket (msm.cu K2b).  One LANE owns a task and reduces its list as a
// binary tree, level by level: level l turns m items into ceil(m / 2) by adding neighbours (2i, 2i + 1) in AFFINE coordinates,
//     lambda = (y2 - y1) / (x2 - x1),  x3 = lambda^2 - x1 - x2,  y3 = lambda (x1 - x3) - y1,
// where the division is shared (Montgomery's trick): in one batch every lane walks up to K pairs of each of its Q tasks, multiplies
// the denominators into a running product (exclusive prefixes parked in a per-warp slab, 1 KB rows = one entry per lane), the 32
// lane products are combined by shuffle scans, ONE field inversion is done per warp and batch (<= 32 Q K additions), and the
// inverse is distributed back along every lane's prefixes: 3 multiplications per denominator + 3 for the formula.
// Warps are independent (no block barrier): while one lane of a warp runs the inversion — integer-ALU work — the other warps of
// the scheduler keep the multiplier busy.  Levels alternate between two scratch regions owned by the task (positions derived
// arithmetically from the task's place in the sorted array, no extra scan).  When fewer than PAIRS_MIN pairs per list remain, the
// rest is folded with the XYZZ mixed addition exactly like k_bucket_accumulate and the task result is written in the same place
// and format, so every later kernel of the pipeline is unchanged.
//
// Exceptional pairs are exact: P + identity, P + P (tangent slope, denominator 2 y), P + (-P) = identity; pairs that need no
// division do not enter the product.  Everything that reaches a shuffle is warp-uniform (derived from the warp-wide maximum
// list length).  Tuning (K, PAIRS_MIN, Q) is in snarkv_ctx (ctx.hpp); measured sweeps: profiles/r01_ba_*.txt.
//
// The kernel is co-limited: 79 % of the integer-multiplier pipe and 3.5 TB/s of DRAM traffic (gathers, tree levels, prefixes) at
// 2^24 terms (profiles/r01_ncu_bucket_accumulate_affine_2p24.txt).  Measured and rejected: prefetch.global.L2 of the next pair's
// operands (adds traffic: 34.3 -> 35.3 .. 37.4 ms), per-lane cp.async staging of the backward pass's operands one iteration ahead
// (33.5 -> 34.5 ms), 5 or 6 resident blocks per SM (spills), K < 64 (more inversions).
#pragma once
#include "g1.cuh"

namespace snarkv {

#define SNARKV_BA_THREADS 128
#define SNARKV_BA_K_MAX 128     // largest K: pairs of one list per batch (with SNARKV_BA_Q_MAX it sizes the prefix slab)


// ONE copy of the Montgomery multiplication for this kernel (a call costs ~12 register moves on top of ~170 instructions): its
// warps sit in different phases (forward pass, backward pass, inversion, tail), and with the multiplication inlined at all ~45
// sites the code outgrows the instruction caches (measured: "no instruction" became the top stall reason).
static __device__ __noinline__ Fq fq_mul_call(const Fq a, const Fq b) { return fp_mul(a, b); }

// acc += (x2, y2) exactly like xyzz_madd (g1.cuh), through the shared multiplication
static __device__ __noinline__ void ba_xyzz_madd(G1Xyzz& acc, const Fq& x2, const Fq& y2) {
    if (xyzz_is_identity(acc)) {
        acc.x = x2; acc.y = y2; acc.zz = fp_one<FQ>(); acc.zzz = fp_one<FQ>();
        return;
    }
    const Fq u2 = fq_mul_call(x2, acc.zz);
    const Fq s2 = fq_mul_call(y2, acc.zzz);
    const Fq p = fp_sub(u2, acc.x);
    const Fq r = fp_sub(s2, acc.y);
    if (fp_is_zero(p)) {
        if (fp_is_zero(r)) acc = xyzz_dbl_affine(x2, y2);
        else acc = xyzz_identity();
        return;
    }
    const Fq pp = fq_mul_call(p, p);
    const Fq ppp = fq_mul_call(p, pp);
    const Fq q = fq_mul_call(acc.x, pp);
    const Fq x3 = fp_sub(fp_sub(fq_mul_call(r, r), ppp), fp_dbl(q));
    acc.y = fp_sub(fq_mul_call(r, fp_sub(q, x3)), fq_mul_call(acc.y, ppp));
    acc.x = x3;
    acc.zz = fq_mul_call(acc.zz, pp);
    acc.zzz = fq_mul_call(acc.zzz, ppp);
}

// plain (coherent) 128-bit loads: the scratch regions are written by this kernel, so the read-only path (__ldg) is not allowed
__device__ __forceinline__ Fq fq_load_rw(const void* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 lo = q[0], hi = q[1];
    Fq r;
    r.v[0] = lo.x; r.v[1] = lo.y; r.v[2] = lo.z; r.v[3] = lo.w;
    r.v[4] = hi.x; r.v[5] = hi.y; r.v[6] = hi.z; r.v[7] = hi.w;
    return r;
}
__device__ __forceinline__ G1Affine g1_affine_load_rw(const uint8_t* base, size_t idx) {
    G1Affine r;
    r.x = fq_load_rw(base + idx * 64);
    r.y = fq_load_rw(base + idx * 64 + 32);
    return r;
}

// item j of the current level: level 0 = sorted reference (index | sign << 31) into the caller's point array, later levels =
// materialised affine points in the task's scratch region
struct BaTask;
struct BaSource {
    const uint8_t* points;    // level 0
    const uint32_t* list;     // level 0
    const uint8_t* region;    // level >= 1
    bool refs;
    __device__ __forceinline__ void make(const BaTask& tk, uint32_t level, const uint8_t* pts, uint8_t* region_a, uint8_t* region_b,
                                         size_t stride_a, size_t stride_b, uint32_t W);
    __device__ __forceinline__ uint8_t* dest(const BaTask& tk, uint32_t level, uint8_t* region_a, uint8_t* region_b, size_t stride_a,
                                             size_t stride_b, uint32_t W) const;
    __device__ __forceinline__ G1Affine get(uint32_t j) const {
        if (refs) {
            const uint32_t e = list[j];
            G1Affine p = g1_affine_load(points, e & 0x7fffffffu);
            if (e >> 31) p.y = fp_neg(p.y);   // -(0, 0) = (0, 0): the identity stays the identity
            return p;
        }
        return g1_affine_load_rw(region, j);
    }
    // Operands of pair i (items 2i, 2i + 1) in two steps, so that the level-0 index loads get a whole iteration to complete:
    // pair_ref requests the two sorted references, pair_addr turns them into addresses + sign bits (bit 0: negate a, bit 1: b).
    __device__ __forceinline__ void pair_ref(uint32_t i, uint32_t& e0, uint32_t& e1) const {
        if (refs) { e0 = list[2 * i]; e1 = list[2 * i + 1]; }
    }
    __device__ __forceinline__ void pair_addr(uint32_t i, uint32_t e0, uint32_t e1, const uint8_t*& pa, const uint8_t*& pb,
                                              uint32_t& signs) const {
        if (refs) {
            pa = points + (size_t)(e0 & 0x7fffffffu) * 64;
            pb = points + (size_t)(e1 & 0x7fffffffu) * 64;
            signs = (e0 >> 31) | ((e1 >> 31) << 1);
        } else {
            pa = region + (size_t)(2 * i) * 64;
            pb = pa + 64;
            signs = 0;
        }
    }
    // one coordinate (32 B): the caller's array goes through the read-only path, this kernel's scratch through coherent loads
    __device__ __forceinline__ Fq load_coord(const uint8_t* p) const { return refs ? fp_load<FQ>(p) : fq_load_rw(p); }
};

// kind of a pair addition a + b: 0 = chord (d = x2 - x1), 1 = tangent (d = 2 y1), 2 = a is the identity (result b),
// 3 = b is the identity (result a), 4 = result is the identity.  Only kinds 0 and 1 need 1 / d.
__device__ __forceinline__ int ba_classify(const G1Affine& a, const G1Affine& b, Fq& d) {
    if (g1_affine_is_identity(a)) return 2;
    if (g1_affine_is_identity(b)) return 3;
    d = fp_sub(b.x, a.x);
    if (!fp_is_zero(d)) return 0;
    if (fp_eq(a.y, b.y) && !fp_is_zero(a.y)) {
        d = fp_dbl(a.y);
        return 1;
    }
    return 4;
}

// ---- warp-wide Montgomery trick -----------------------------------------------------------------------------------------
__device__ __forceinline__ Fq fq_shfl(const Fq& a, int src_lane) {
    Fq r;
#pragma unroll
    for (int k = 0; k < 8; ++k) r.v[k] = __shfl_sync(0xffffffffu, a.v[k], src_lane);
    return r;
}
__device__ __forceinline__ Fq fq_shfl_up(const Fq& a, uint32_t delta) {
    Fq r;
#pragma unroll
    for (int k = 0; k < 8; ++k) r.v[k] = __shfl_up_sync(0xffffffffu, a.v[k], delta);
    return r;
}
__device__ __forceinline__ Fq fq_shfl_down(const Fq& a, uint32_t delta) {
    Fq r;
#pragma unroll
    for (int k = 0; k < 8; ++k) r.v[k] = __shfl_down_sync(0xffffffffu, a.v[k], delta);
    return r;
}
__device__ __forceinline__ Fq fq_select(bool c, const Fq& a, const Fq& b) {
    Fq r;
#pragma unroll
    for (int k = 0; k < 8; ++k) r.v[k] = c ? a.v[k] : b.v[k];
    return r;
}
// Every lane passes a non-zero `run`; returns 1 / run with ONE field inversion for the warp: inclusive prefix and suffix
// products by shuffle scans (5 steps each), lane 0 inverts the total, 1 / run_l = inv * (prefix before l) * (suffix after l).
// The inversion is integer-ALU work of one lane; the other warps of the scheduler keep the multiplier busy meanwhile.
static __device__ __noinline__ Fq ba_warp_invert(const Fq& run, uint32_t lane) {
    __syncwarp();
    Fq pre = run, suf = run;
#pragma unroll 1
    for (uint32_t o = 1; o < 32; o <<= 1) {
        const Fq mp = fq_mul_call(pre, fq_shfl_up(pre, o));
        const Fq ms = fq_mul_call(suf, fq_shfl_down(suf, o));
        pre = fq_select(lane >= o, mp, pre);
        suf = fq_select(lane + o < 32, ms, suf);
    }
    Fq inv = fq_shfl(pre, 31);
    if (lane == 0) inv = fp_inv_serial(inv);
    __syncwarp();
    inv = fq_shfl(inv, 0);
    const Fq before = fq_select(lane > 0, fq_shfl_up(pre, 1), fp_one<FQ>());
    const Fq after = fq_select(lane < 31, fq_shfl_down(suf, 1), fp_one<FQ>());
    return fq_mul_call(fq_mul_call(inv, before), after);
}

// per-task state, parked in shared memory between the phases of a group (one entry per lane and task slot q)
struct BaTask {
    const uint32_t* list;   // the task's run of sorted references
    uint32_t sa, sb;        // first item of its level-1/3/.. and level-2/4/.. scratch regions (within window w of base set z)
    uint32_t m0;            // list length; 0 = no task
    uint32_t slot;          // task slot within the window (where the result goes)
    uint32_t w, z;
};

#define SNARKV_BA_Q_MAX 4   // tasks per lane that share one inversion

// level 0 reads the sorted references; level l >= 1 reads what level l - 1 wrote: region A after even levels, B after odd ones
__device__ __forceinline__ void BaSource::make(const BaTask& tk, uint32_t level, const uint8_t* pts, uint8_t* region_a, uint8_t* region_b,
                                               size_t stride_a, size_t stride_b, uint32_t W) {
    points = pts;
    list = tk.list;
    refs = level == 0;
    const size_t win = (size_t)tk.z * W + tk.w;
    region = (level & 1u) ? region_a + (win * stride_a + tk.sa) * 64 : region_b + (win * stride_b + tk.sb) * 64;
}
__device__ __forceinline__ uint8_t* BaSource::dest(const BaTask& tk, uint32_t level, uint8_t* region_a, uint8_t* region_b, size_t stride_a,
                                                   size_t stride_b, uint32_t W) const {
    const size_t win = (size_t)tk.z * W + tk.w;
    return (level & 1u) ? region_b + (win * stride_b + tk.sb) * 64 : region_a + (win * stride_a + tk.sa) * 64;
}

// Persistent kernel of independent warps.  Work is handed out in units of 32 length-ordered tasks (unit u -> rank group
// gi = u / (W Z), then base set z and window w, so that the longest tasks of all windows come first); a warp takes Q units at a
// time (fewer near the end of the queue, for balance) and runs their 32 Q lists level by level, one shared inversion per batch
// of up to K pairs of each list.
__global__ void __launch_bounds__(SNARKV_BA_THREADS, 4)
k_bucket_accumulate_affine(const uint8_t* __restrict__ points0, const uint8_t* __restrict__ points1, const uint32_t* __restrict__ sorted,
                           const uint32_t* __restrict__ offsets, const uint32_t* __restrict__ counts, const uint2* __restrict__ tasks,
                           const uint32_t* __restrict__ window_tasks, const uint32_t* __restrict__ order, size_t n, uint32_t NB,
                           uint32_t T, uint32_t cap, uint32_t W, uint32_t Z, uint8_t* __restrict__ task_out, uint8_t* region_a,
                           uint8_t* region_b, size_t region_a_stride, size_t region_b_stride, uint8_t* prefix_slab,
                           uint32_t* group_counter, uint32_t K, uint32_t pairs_min, uint32_t Qmax) {
    __shared__ BaTask s_task[SNARKV_BA_Q_MAX][SNARKV_BA_THREADS];
    const uint32_t t = threadIdx.x, lane = t & 31u;
    const uint32_t warp_global = (blockIdx.x * SNARKV_BA_THREADS + t) >> 5, n_warps = (gridDim.x * SNARKV_BA_THREADS) >> 5;
    // units beyond the longest window's task count are empty: the queue ends at gi_end
    uint32_t wt_max = 0;
    for (uint32_t w = lane; w < W; w += 32) wt_max = max(wt_max, window_tasks[w]);
    wt_max = __reduce_max_sync(0xffffffffu, wt_max);
    const uint32_t total_units = ((wt_max + 31u) >> 5) * W * Z;
    uint8_t* pref = prefix_slab + ((size_t)warp_global * Qmax * K * 32 + lane) * 32;   // Qmax K rows of 32 lanes x 32 B
    for (;;) {
        // take Q units: Qmax while the queue holds that many per warp, fewer in the last round (balance)
        uint32_t u0 = 0, Q = 0;
        if (lane == 0) {
            const uint32_t seen = *(volatile uint32_t*)group_counter;
            const uint32_t left = seen < total_units ? total_units - seen : 0u;
            Q = min(Qmax, max(1u, left / n_warps));
            u0 = atomicAdd(group_counter, Q);
        }
        u0 = __shfl_sync(0xffffffffu, u0, 0);
        Q = __shfl_sync(0xffffffffu, Q, 0);
        if (u0 >= total_units) break;
        // ---- task set-up ----
        uint32_t mmax = 0;
        for (uint32_t q = 0; q < Q; ++q) {
            BaTask tk;
            tk.list = sorted; tk.sa = tk.sb = tk.m0 = tk.slot = tk.w = tk.z = 0;
            const uint32_t u = u0 + q;
            if (u < total_units) {
                const uint32_t gi = u / (W * Z), rem = u - gi * (W * Z);
                tk.z = rem / W;
                tk.w = rem - tk.z * W;
                const uint32_t rank = gi * 32u + lane;
                if (rank < window_tasks[tk.w]) {
                    tk.slot = order[(size_t)tk.w * cap + rank];
                    const uint2 task = tasks[(size_t)tk.w * cap + tk.slot];
                    const uint32_t bucket = tk.w * NB + task.x;
                    const uint32_t first = task.y * T;
                    tk.m0 = min(T, counts[bucket] - first);
                    const uint32_t pos = offsets[bucket] + first;
                    tk.list = sorted + (size_t)tk.w * n + pos;
                    // level-1 region: items [sa, sa + ceil(m/2)); level-2 region: [sb, sb + ceil(ceil(m/2)/2)).  Consecutive tasks
                    // (pos' = pos + m, slot' = slot + 1) get disjoint regions: floor((X + m + 1) / 2) - floor(X / 2) >= ceil(m / 2).
                    tk.sa = (pos + tk.slot + 1u) >> 1;
                    tk.sb = (tk.sa + tk.slot + 1u) >> 1;
                }
            }
            s_task[q][t] = tk;
            mmax = max(mmax, tk.m0);
        }
        mmax = __reduce_max_sync(0xffffffffu, mmax);

        uint32_t level = 0;
#pragma unroll 1
        for (;; ++level) {
            const uint32_t mmax_l = (mmax + (1u << level) - 1u) >> level;
            const uint32_t maxpairs = mmax_l >> 1;
            if (maxpairs < pairs_min || level >= 31) break;
#pragma unroll 1
            for (uint32_t cb = 0; cb < maxpairs; cb += K) {
                // forward: exclusive prefix products of the denominators.  Only the x coordinates are needed unless the pair is
                // exceptional; the loads of pair i + 1 are issued before the multiplication of pair i, its references one earlier.
                Fq run = fp_one<FQ>();
#pragma unroll 1
                for (uint32_t q = 0; q < Q; ++q) {
                    const BaTask tk = s_task[q][t];
                    const uint32_t m = (tk.m0 + (1u << level) - 1u) >> level, pairs = m >> 1;
                    const uint32_t lo = min(cb, pairs), hi = min(cb + K, pairs);
                    BaSource src;
                    src.make(tk, level, tk.z == 0 ? points0 : points1, region_a, region_b, region_a_stride, region_b_stride, W);
                    uint8_t* pq = pref + (size_t)q * K * 1024;
                    const uint8_t *pa = nullptr, *pb = nullptr;
                    uint32_t sg = 0, e0 = 0, e1 = 0;
                    Fq axn = fp_zero<FQ>(), bxn = fp_zero<FQ>();
                    if (lo < hi) {
                        src.pair_ref(lo, e0, e1);
                        src.pair_addr(lo, e0, e1, pa, pb, sg);
                        axn = src.load_coord(pa);
                        bxn = src.load_coord(pb);
                        if (lo + 1 < hi) src.pair_ref(lo + 1, e0, e1);
                    }
#pragma unroll 1
                    for (uint32_t i = lo; i < hi; ++i) {
                        const Fq ax = axn, bx = bxn;
                        if (i + 1 < hi) {
                            src.pair_addr(i + 1, e0, e1, pa, pb, sg);
                            axn = src.load_coord(pa);
                            bxn = src.load_coord(pb);
                            if (i + 2 < hi) src.pair_ref(i + 2, e0, e1);
                        }
                        Fq d = fp_sub(bx, ax);
                        if (fp_is_zero(ax) || fp_is_zero(bx) || fp_is_zero(d)) {   // rare: identity operand, equal or opposite points
                            const G1Affine a = src.get(2 * i), b = src.get(2 * i + 1);
                            if (ba_classify(a, b, d) > 1) continue;
                        }
                        fp_store<FQ>(pq + (size_t)(i - lo) * 1024, run);
                        run = fp_mul(run, d);   // inlined: a call would first wait for the loads of pair i + 1 issued above
                    }
                }
                Fq acc = ba_warp_invert(run, lane);
                // backward: 1 / d_i = acc * prefix_i, acc *= d_i; then the chord / tangent formula.  The four coordinates of a pair
                // are requested at the top of its iteration, its prefix and its sorted references one iteration earlier, so that
                // the multiplication acc * prefix_i runs while the gathered coordinates are still in flight.
#pragma unroll 1
                for (uint32_t q = Q; q-- > 0;) {
                    const BaTask tk = s_task[q][t];
                    const uint32_t m = (tk.m0 + (1u << level) - 1u) >> level, pairs = m >> 1;
                    const uint32_t lo = min(cb, pairs), hi = min(cb + K, pairs);
                    BaSource src;
                    src.make(tk, level, tk.z == 0 ? points0 : points1, region_a, region_b, region_a_stride, region_b_stride, W);
                    uint8_t* dst = src.dest(tk, level, region_a, region_b, region_a_stride, region_b_stride, W);
                    const uint8_t* pq = pref + (size_t)q * K * 1024;
                    const uint8_t *pa = nullptr, *pb = nullptr;
                    uint32_t sg = 0, e0 = 0, e1 = 0;
                    Fq pfn = fp_zero<FQ>();
                    if (lo < hi) {
                        src.pair_ref(hi - 1, e0, e1);
                        pfn = fq_load_rw(pq + (size_t)(hi - 1 - lo) * 1024);
                    }
#pragma unroll 1
                    for (uint32_t i = hi; i-- > lo;) {
                        const Fq pf = pfn;
                        src.pair_addr(i, e0, e1, pa, pb, sg);
                        G1Affine a, b;
                        a.x = src.load_coord(pa); a.y = src.load_coord(pa + 32);
                        b.x = src.load_coord(pb); b.y = src.load_coord(pb + 32);
                        if (i > lo) {
                            src.pair_ref(i - 1, e0, e1);
                            pfn = fq_load_rw(pq + (size_t)(i - 1 - lo) * 1024);
                        }
                        // inlined (a call would first wait for the coordinate loads issued above); unused, and pf undefined, for
                        // pairs that need no division
                        const Fq inv = fp_mul(acc, pf);
                        if (sg & 1u) a.y = fp_neg(a.y);   // level 0 only (the loads above are already in flight)
                        if (sg >> 1) b.y = fp_neg(b.y);
                        Fq d;
                        const int kind = ba_classify(a, b, d);
                        G1Affine o;
                        if (kind <= 1) {
                            acc = fq_mul_call(acc, d);
                            Fq num;
                            if (kind == 0) num = fp_sub(b.y, a.y);
                            else {
                                const Fq xx = fq_mul_call(a.x, a.x);
                                num = fp_add(fp_dbl(xx), xx);
                            }
                            const Fq lam = fq_mul_call(num, inv);
                            o.x = fp_sub(fp_sub(fq_mul_call(lam, lam), a.x), b.x);
                            o.y = fp_sub(fq_mul_call(lam, fp_sub(a.x, o.x)), a.y);
                        } else if (kind == 2) o = b;
                        else if (kind == 3) o = a;
                        else { o.x = fp_zero<FQ>(); o.y = fp_zero<FQ>(); }
                        g1_affine_store(dst, i, o);
                    }
                }
            }
            // an odd item out moves up unchanged
            for (uint32_t q = 0; q < Q; ++q) {
                const BaTask tk = s_task[q][t];
                const uint32_t m = (tk.m0 + (1u << level) - 1u) >> level;
                if (m & 1u) {
                    BaSource src;
                    src.make(tk, level, tk.z == 0 ? points0 : points1, region_a, region_b, region_a_stride, region_b_stride, W);
                    g1_affine_store(src.dest(tk, level, region_a, region_b, region_a_stride, region_b_stride, W), m >> 1, src.get(m - 1));
                }
            }
        }
        // tail: fold what is left with the XYZZ mixed addition (k_bucket_accumulate's loop) and emit the task results
#pragma unroll 1
        for (uint32_t q = 0; q < Q; ++q) {
            const BaTask tk = s_task[q][t];
            if (tk.m0 == 0) continue;
            const uint32_t m = (tk.m0 + (1u << level) - 1u) >> level;
            BaSource src;
            src.make(tk, level, tk.z == 0 ? points0 : points1, region_a, region_b, region_a_stride, region_b_stride, W);
            G1Xyzz acc = xyzz_identity();
#pragma unroll 1
            for (uint32_t k = 0; k < m; ++k) {
                const G1Affine cur = src.get(k);
                if (g1_affine_is_identity(cur)) continue;
                ba_xyzz_madd(acc, cur.x, cur.y);
            }
            xyzz_store(task_out + (size_t)tk.z * W * cap * 128, (size_t)tk.w * cap + tk.slot, acc);
        }
        __syncwarp();
    }
}

// Self-check support (accumulate mode 3): compares two task-result arrays as group elements.
__global__ void __launch_bounds__(128) k_compare_task_results(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b,
                                                              const uint32_t* __restrict__ window_tasks, const uint32_t* __restrict__ order,
                                                              uint32_t cap, uint32_t W, uint32_t* __restrict__ mismatch) {
    const uint32_t w = blockIdx.y, z = blockIdx.z;
    const uint32_t rank = blockIdx.x * blockDim.x + threadIdx.x;
    if (rank >= window_tasks[w]) return;
    const uint32_t slot = order[(size_t)w * cap + rank];
    const size_t idx = ((size_t)z * W + w) * cap + slot;
    const G1Xyzz p = xyzz_load(a, idx), q = xyzz_load(b, idx);
    bool same;
    if (xyzz_is_identity(p) || xyzz_is_identity(q)) same = xyzz_is_identity(p) && xyzz_is_identity(q);
    else same = fp_eq(fq_mul_call(p.x, q.zz), fq_mul_call(q.x, p.zz)) && fp_eq(fq_mul_call(p.y, q.zzz), fq_mul_call(q.y, p.zzz));
    if (!same) {
        const uint32_t k = atomicAdd(&mismatch[0], 1u);
        if (k == 0) { mismatch[1] = w; mismatch[2] = slot; mismatch[3] = z; }
    }
}

}  // namespace snarkv
