// bucket_affine.cuh — batched-affine bucket accumulation: the same `buckets[scalar - 1].add_assign(base)` of
// snark-verifier/src/util/msm.rs:291-296 as k_bucket_accumulate (msm.cu), with 6 instead of 10 Montgomery multiplications
// per point addition.
//
// A task is a run of <= T sorted point references of one bucket (msm.cu K2b).  One thread owns a task and reduces its list as a
// binary tree, level by level: level l turns m items into ceil(m / 2) by adding neighbours (2i, 2i + 1) in AFFINE coordinates,
//     lambda = (y2 - y1) / (x2 - x1),  x3 = lambda^2 - x1 - x2,  y3 = lambda (x1 - x3) - y1,
// where the division is shared: the 128 threads of a block each run K pair-additions per batch, multiply their K denominators
// into a running product (exclusive prefixes parked in an L2-resident slab), the 128 thread products are combined in a
// shared-memory product tree, ONE field inversion is done per batch (<= 128 K additions), and the inverse is distributed back
// down the tree and along every thread's prefixes (Montgomery's trick): 3 multiplications per denominator + 3 for the formula.
// Levels alternate between two scratch regions owned by the task (positions derived arithmetically from the task's place in
// the sorted array, no extra scan).  When fewer than PAIRS_MIN pairs per thread remain, the rest of the list is folded with the
// XYZZ mixed addition exactly like k_bucket_accumulate and the task result is written in the same place and format, so every
// later kernel of the pipeline is unchanged.
//
// Exceptional pairs are exact: P + identity, P + P (tangent slope, denominator 2 y), P + (-P) = identity; such pairs that need
// no division do not enter the product.  All control flow that reaches a __syncthreads is block-uniform (derived from the
// block-wide maximum list length).
#pragma once
#include "g1.cuh"

namespace snarkv {

#define SNARKV_BA_THREADS 128
#define SNARKV_BA_K_MAX 64      // largest batch: pair-additions per thread per shared inversion (sizes the prefix slab)
#define SNARKV_BA_K 32          // default batch
#define SNARKV_BA_PAIRS_MIN 12  // default: run another affine level while the longest list of the block still has this many pairs


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
struct BaSource {
    const uint8_t* points;    // level 0
    const uint32_t* list;     // level 0
    const uint8_t* region;    // level >= 1
    bool refs;
    __device__ __forceinline__ G1Affine get(uint32_t j) const {
        if (refs) {
            const uint32_t e = list[j];
            G1Affine p = g1_affine_load(points, e & 0x7fffffffu);
            if (e >> 31) p.y = fp_neg(p.y);   // -(0, 0) = (0, 0): the identity stays the identity
            return p;
        }
        return g1_affine_load_rw(region, j);
    }
    // addresses of the two operands of pair i (items 2i, 2i + 1) and their sign bits (bit 0: negate a, bit 1: negate b)
    __device__ __forceinline__ void pair_addr(uint32_t i, const uint8_t*& pa, const uint8_t*& pb, uint32_t& signs) const {
        if (refs) {
            const uint32_t e0 = list[2 * i], e1 = list[2 * i + 1];
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

// branch-free conditional negation (the point loads of a pair stay back to back instead of being split by a branch)
__device__ __forceinline__ Fq fq_cneg(const Fq& y, uint32_t neg) {
    const Fq ny = fp_neg(y);
    const uint32_t mask = 0u - neg;
    Fq r;
#pragma unroll
    for (int k = 0; k < 8; ++k) r.v[k] = (ny.v[k] & mask) | (y.v[k] & ~mask);
    return r;
}

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

// tree[1] = product of the 128 leaves tree[128 + t]; afterwards tree[128 + t] = 1 / leaf_t.  Heap layout, in place.
// Called by all SNARKV_BA_THREADS threads; `inverter` is the thread that performs the single field inversion.
__device__ __forceinline__ void ba_block_invert(Fq* tree, uint32_t t, uint32_t inverter) {
    __syncthreads();
#pragma unroll 1
    for (uint32_t s = SNARKV_BA_THREADS / 2; s >= 1; s >>= 1) {
        if (t < s) tree[s + t] = fp_mul(tree[2 * (s + t)], tree[2 * (s + t) + 1]);
        __syncthreads();
    }
    if (t == inverter) tree[1] = fp_inv_serial(tree[1]);
    __syncthreads();
#pragma unroll 1
    for (uint32_t s = 1; s < SNARKV_BA_THREADS; s <<= 1) {
        if (t < s) {
            const uint32_t i = s + t;
            const Fq inv = tree[i], l = tree[2 * i], r = tree[2 * i + 1];
            tree[2 * i] = fp_mul(inv, r);
            tree[2 * i + 1] = fp_mul(inv, l);
        }
        __syncthreads();
    }
}

// Persistent kernel: blocks pull groups of 128 length-ordered tasks from a global counter.  Group g -> (rank group gi, window
// w, base set z) with the rank group outermost, so the longest tasks of all windows are started first.
__global__ void __launch_bounds__(SNARKV_BA_THREADS, 4)
k_bucket_accumulate_affine(const uint8_t* __restrict__ points0, const uint8_t* __restrict__ points1, const uint32_t* __restrict__ sorted,
                           const uint32_t* __restrict__ offsets, const uint32_t* __restrict__ counts, const uint2* __restrict__ tasks,
                           const uint32_t* __restrict__ window_tasks, const uint32_t* __restrict__ order, size_t n, uint32_t NB,
                           uint32_t T, uint32_t cap, uint32_t W, uint32_t Z, uint8_t* __restrict__ task_out, uint8_t* region_a,
                           uint8_t* region_b, size_t region_a_stride, size_t region_b_stride, uint8_t* prefix_slab,
                           uint32_t* __restrict__ group_counter, uint32_t K, uint32_t pairs_min) {
    __shared__ Fq tree[2 * SNARKV_BA_THREADS];
    __shared__ uint32_t s_group, s_max[SNARKV_BA_THREADS / 32];
    const uint32_t t = threadIdx.x;
    const uint32_t groups_per_window = (cap + SNARKV_BA_THREADS - 1) / SNARKV_BA_THREADS;
    const uint32_t total_groups = groups_per_window * W * Z;
    uint8_t* pref = prefix_slab + (size_t)blockIdx.x * SNARKV_BA_K_MAX * SNARKV_BA_THREADS * 32 + (size_t)t * 32;
    uint32_t batch_no = blockIdx.x;   // rotates the inverting thread over the four schedulers
    for (;;) {
        __syncthreads();
        if (t == 0) s_group = atomicAdd(group_counter, 1u);
        __syncthreads();
        const uint32_t g = s_group;
        if (g >= total_groups) break;
        const uint32_t gi = g / (W * Z), rem = g - gi * (W * Z);
        const uint32_t z = rem / W, w = rem - z * W;
        const uint32_t rank = gi * SNARKV_BA_THREADS + t;
        if (gi * SNARKV_BA_THREADS >= window_tasks[w]) continue;   // block-uniform
        const bool active = rank < window_tasks[w];
        const uint8_t* __restrict__ points = z == 0 ? points0 : points1;
        uint32_t slot = 0, m = 0;
        BaSource src;
        src.points = points; src.list = sorted; src.region = nullptr; src.refs = true;
        uint8_t *reg_a = nullptr, *reg_b = nullptr;
        if (active) {
            slot = order[(size_t)w * cap + rank];
            const uint2 task = tasks[(size_t)w * cap + slot];
            const uint32_t bucket = w * NB + task.x;
            const uint32_t first = task.y * T;
            m = min(T, counts[bucket] - first);
            const uint32_t pos = offsets[bucket] + first;
            src.list = sorted + (size_t)w * n + pos;
            // level-1 region: items [sa, sa + ceil(m/2)); level-2 region: [sb, sb + ceil(ceil(m/2)/2)).  Consecutive tasks
            // (pos' = pos + m, slot' = slot + 1) get disjoint regions: floor((X + m + 1) / 2) - floor(X / 2) >= ceil(m / 2).
            const uint32_t sa = (pos + slot + 1u) >> 1;
            const uint32_t sb = (sa + slot + 1u) >> 1;
            reg_a = region_a + ((size_t)(z * W + w) * region_a_stride + sa) * 64;
            reg_b = region_b + ((size_t)(z * W + w) * region_b_stride + sb) * 64;
        }
        // block-wide maximum list length
        uint32_t mmax = __reduce_max_sync(0xffffffffu, m);
        if ((t & 31u) == 0) s_max[t >> 5] = mmax;
        __syncthreads();
        mmax = max(max(s_max[0], s_max[1]), max(s_max[2], s_max[3]));

        uint32_t level = 0;
#pragma unroll 1
        while ((mmax >> 1) >= pairs_min) {
            const uint32_t pairs = m >> 1, maxpairs = mmax >> 1;
            uint8_t* dst = (level & 1u) ? reg_b : reg_a;
#pragma unroll 1
            for (uint32_t cb = 0; cb < maxpairs; cb += K, ++batch_no) {
                const uint32_t lo = min(cb, pairs), hi = min(cb + K, pairs);
                // forward: exclusive prefix products of the denominators.  Only the x coordinates are needed unless the pair is
                // exceptional; the loads of pair i + 1 are issued before the multiplication of pair i, its addresses one pair earlier.
                Fq run = fp_one<FQ>();
                const uint8_t *pa_n = nullptr, *pb_n = nullptr;
                uint32_t sg_n = 0;
                Fq axn = fp_zero<FQ>(), bxn = fp_zero<FQ>();
                if (lo < hi) {
                    src.pair_addr(lo, pa_n, pb_n, sg_n);
                    axn = src.load_coord(pa_n);
                    bxn = src.load_coord(pb_n);
                    if (lo + 1 < hi) src.pair_addr(lo + 1, pa_n, pb_n, sg_n);
                }
#pragma unroll 1
                for (uint32_t i = lo; i < hi; ++i) {
                    const Fq ax = axn, bx = bxn;
                    if (i + 1 < hi) {
                        axn = src.load_coord(pa_n);
                        bxn = src.load_coord(pb_n);
                        if (i + 2 < hi) src.pair_addr(i + 2, pa_n, pb_n, sg_n);
                    }
                    Fq d = fp_sub(bx, ax);
                    if (fp_is_zero(ax) || fp_is_zero(bx) || fp_is_zero(d)) {   // rare: identity operand, equal or opposite points
                        const G1Affine a = src.get(2 * i), b = src.get(2 * i + 1);
                        if (ba_classify(a, b, d) > 1) continue;
                    }
                    fp_store<FQ>(pref + (size_t)(i - lo) * SNARKV_BA_THREADS * 32, run);
                    run = fp_mul(run, d);
                }
                tree[SNARKV_BA_THREADS + t] = run;
                ba_block_invert(tree, t, (batch_no & 3u) * 32u);
                Fq acc = tree[SNARKV_BA_THREADS + t];
                // backward: 1 / d_i = acc * prefix_i, acc *= d_i; then the chord / tangent formula.  All five operands of a pair
                // are requested at the top of its iteration and the prefix (L2-resident slab) is consumed first, so that the
                // multiplication acc * prefix_i runs while the gathered coordinates are still in flight.
                if (lo < hi) src.pair_addr(hi - 1, pa_n, pb_n, sg_n);
#pragma unroll 1
                for (uint32_t i = hi; i-- > lo;) {
                    const Fq pf = fq_load_rw(pref + (size_t)(i - lo) * SNARKV_BA_THREADS * 32);
                    G1Affine a, b;
                    a.x = src.load_coord(pa_n); a.y = src.load_coord(pa_n + 32);
                    b.x = src.load_coord(pb_n); b.y = src.load_coord(pb_n + 32);
                    const uint32_t sg = sg_n;
                    if (i > lo) src.pair_addr(i - 1, pa_n, pb_n, sg_n);
                    const Fq inv = fp_mul(acc, pf);          // unused (and pf undefined) for pairs that need no division
                    a.y = fq_cneg(a.y, sg & 1u);
                    b.y = fq_cneg(b.y, sg >> 1);
                    Fq d;
                    const int kind = ba_classify(a, b, d);
                    G1Affine o;
                    if (kind <= 1) {
                        acc = fp_mul(acc, d);
                        Fq num;
                        if (kind == 0) num = fp_sub(b.y, a.y);
                        else {
                            const Fq xx = fp_sqr(a.x);
                            num = fp_add(fp_dbl(xx), xx);
                        }
                        const Fq lam = fp_mul(num, inv);
                        o.x = fp_sub(fp_sub(fp_sqr(lam), a.x), b.x);
                        o.y = fp_sub(fp_mul(lam, fp_sub(a.x, o.x)), a.y);
                    } else if (kind == 2) o = b;
                    else if (kind == 3) o = a;
                    else { o.x = fp_zero<FQ>(); o.y = fp_zero<FQ>(); }
                    g1_affine_store(dst, i, o);
                }
            }
            if (m & 1u) g1_affine_store(dst, pairs, src.get(m - 1));
            if (active) { src.refs = false; src.region = dst; }
            m = pairs + (m & 1u);
            mmax = (mmax >> 1) + (mmax & 1u);
            ++level;
        }
        // tail: fold what is left with the XYZZ mixed addition (k_bucket_accumulate's loop) and emit the task result
        if (active) {
            G1Xyzz acc = xyzz_identity();
#pragma unroll 1
            for (uint32_t k = 0; k < m; ++k) {
                const G1Affine cur = src.get(k);
                if (g1_affine_is_identity(cur)) continue;
                xyzz_madd(acc, cur.x, cur.y);
            }
            xyzz_store(task_out + (size_t)z * W * cap * 128, (size_t)w * cap + slot, acc);
        }
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
    else same = fp_eq(fp_mul(p.x, q.zz), fp_mul(q.x, p.zz)) && fp_eq(fp_mul(p.y, q.zzz), fp_mul(q.y, p.zzz));
    if (!same) {
        const uint32_t k = atomicAdd(&mismatch[0], 1u);
        if (k == 0) { mismatch[1] = w; mismatch[2] = slot; mismatch[3] = z; }
    }
}

}  // namespace snarkv
