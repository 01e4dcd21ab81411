// bucket_chain.cuh — chained batched-affine bucket accumulation: `buckets[scalar - 1].add_assign(base)` of
// snark-verifier/src/util/msm.rs:291-296 with 6 Montgomery multiplications per point addition, every input point read from HBM
// once, and no intermediate point ever written to HBM.
//
// Why a second affine kernel: k_bucket_accumulate_affine (bucket_affine.cuh) reduces a list as a binary tree.  Its batches are
// 32 x Q x K additions wide, so the inversion prefixes (1.2 GB) and every tree level go through HBM: 120 GB of DRAM traffic per
// 2^24-term launch against 1.6 GB of input (r01 ncu capture), and the last < 24 pairs of every list fall back to XYZZ additions.
// Here a LANE owns a task (a run of <= T sorted references of one bucket) and keeps R running sums ("chains"): item i of the list
// belongs to chain i mod R.  One step adds the next R items to the R chains — R independent affine additions that share ONE
// inversion (Montgomery's trick) done by the lane itself (binary GCD, fp_inv.cuh: integer-ALU work, while the other warps of the
// scheduler keep the multiplier busy).  What this buys:
//   * the R prefix products of a step live in SHARED memory (R x 32 B per lane), not in a global slab;
//   * the R running sums of a lane live in a per-warp slab of R x 2 KB (coalesced 16-byte accesses: [chain][quarter][lane]) that
//     is re-read one step later, i.e. out of L2 — HBM sees only the gathers of the input points;
//   * the backward pass of step s (finish the additions) and the forward pass of step s + 1 (denominators of the next additions)
//     are ONE loop: iteration r finishes chain r's addition and, with the sum still in registers, multiplies the next
//     denominator into the running product.  Products commute, so the chains are simply visited in alternating directions;
//   * lanes never communicate (no shuffles, no warp-wide inversion), lists are affine down to their last R items; the R chain sums
//     are folded with the XYZZ mixed addition and written exactly where k_bucket_accumulate writes, so the rest of the pipeline is
//     unchanged.
// Exceptional pairs are exact as in bucket_affine.cuh: P + O, O + P, P + P (tangent), P + (-P).
#pragma once
#include "bucket_affine.cuh"

namespace snarkv {

#define SNARKV_BC_THREADS 128

struct BcOperands {   // what one loop iteration consumes; requested one iteration ahead
    Fq ax, ay;        // chain sum (steps >= 1)
    Fq bx, by;        // the item added in this step
    Fq xn;            // x of the item this chain receives in the NEXT step
    uint32_t eb, en;  // their sorted references
};

template <int R>
struct BcLane {
    const uint8_t* points;
    const uint32_t* list;
    uint32_t m;
    uint4* slab;          // this lane's column of the warp's chain-sum slab: entry (r, quarter) at slab[(r * 4 + quarter) * 32]
    uint4* pre;           // this lane's column of the warp's prefix rows: entry (r, half) at pre[(r * 2 + half) * 32]

    __device__ __forceinline__ Fq slab_load(uint32_t r, uint32_t half) const {   // coherent: written by this kernel
        const uint4 lo = slab[(r * 4 + half * 2) * 32], hi = slab[(r * 4 + half * 2 + 1) * 32];
        Fq v;
        v.v[0] = lo.x; v.v[1] = lo.y; v.v[2] = lo.z; v.v[3] = lo.w;
        v.v[4] = hi.x; v.v[5] = hi.y; v.v[6] = hi.z; v.v[7] = hi.w;
        return v;
    }
    __device__ __forceinline__ void slab_store(uint32_t r, const G1Affine& p) const {
        slab[(r * 4 + 0) * 32] = make_uint4(p.x.v[0], p.x.v[1], p.x.v[2], p.x.v[3]);
        slab[(r * 4 + 1) * 32] = make_uint4(p.x.v[4], p.x.v[5], p.x.v[6], p.x.v[7]);
        slab[(r * 4 + 2) * 32] = make_uint4(p.y.v[0], p.y.v[1], p.y.v[2], p.y.v[3]);
        slab[(r * 4 + 3) * 32] = make_uint4(p.y.v[4], p.y.v[5], p.y.v[6], p.y.v[7]);
    }
    __device__ __forceinline__ Fq pre_load(uint32_t r) const {
        const uint4 lo = pre[(r * 2) * 32], hi = pre[(r * 2 + 1) * 32];
        Fq v;
        v.v[0] = lo.x; v.v[1] = lo.y; v.v[2] = lo.z; v.v[3] = lo.w;
        v.v[4] = hi.x; v.v[5] = hi.y; v.v[6] = hi.z; v.v[7] = hi.w;
        return v;
    }
    __device__ __forceinline__ void pre_store(uint32_t r, const Fq& v) const {
        pre[(r * 2) * 32] = make_uint4(v.v[0], v.v[1], v.v[2], v.v[3]);
        pre[(r * 2 + 1) * 32] = make_uint4(v.v[4], v.v[5], v.v[6], v.v[7]);
    }
    // operands of iteration (s, r): issued as plain loads, consumed one iteration later
    __device__ __forceinline__ void request(uint32_t s, uint32_t r, BcOperands& o) const {
        const uint32_t idx = s * R + r;
        if (idx >= m) return;
        o.eb = list[idx];
        const uint8_t* pb = points + (size_t)(o.eb & 0x7fffffffu) * 64;
        o.bx = fp_load<FQ>(pb);
        o.by = fp_load<FQ>(pb + 32);
        if (s > 0) {
            o.ax = slab_load(r, 0);
            o.ay = slab_load(r, 1);
        }
        if (idx + R < m) {
            o.en = list[idx + R];
            o.xn = fp_load<FQ>(points + (size_t)(o.en & 0x7fffffffu) * 64);
        }
    }
};

__device__ __forceinline__ G1Affine bc_item(const uint8_t* points, uint32_t e) {
    G1Affine p = g1_affine_load(points, e & 0x7fffffffu);
    if (e >> 31) p.y = fp_neg(p.y);
    return p;
}

// Persistent kernel of independent lanes.  Work is handed out in units of 32 length-ordered tasks exactly like
// k_bucket_accumulate_affine (unit u -> rank group, base set z, window w: the longest tasks of all windows come first).
// Dynamic shared memory: (SNARKV_BC_THREADS / 32) warps x R rows x 1 KB of prefix products.
template <int R>
__global__ void __launch_bounds__(SNARKV_BC_THREADS, 3)
k_bucket_accumulate_chain(const uint8_t* __restrict__ points0, const uint8_t* __restrict__ points1, const uint32_t* __restrict__ sorted,
                          const uint32_t* __restrict__ offsets, const uint32_t* __restrict__ counts, const uint2* __restrict__ tasks,
                          const uint32_t* __restrict__ window_tasks, const uint32_t* __restrict__ order, size_t n, uint32_t NB, uint32_t T,
                          uint32_t cap, uint32_t W, uint32_t Z, uint8_t* __restrict__ task_out, uint4* chain_slab, uint32_t* group_counter) {
    extern __shared__ __align__(16) uint4 bc_prefix[];
    const uint32_t t = threadIdx.x, lane = t & 31u, warp = t >> 5;
    const uint32_t warp_global = (blockIdx.x * SNARKV_BC_THREADS + t) >> 5;
    uint32_t wt_max = 0;
    for (uint32_t w = lane; w < W; w += 32) wt_max = max(wt_max, window_tasks[w]);
    wt_max = __reduce_max_sync(0xffffffffu, wt_max);
    const uint32_t total_units = ((wt_max + 31u) >> 5) * W * Z;
    BcLane<R> ln;
    ln.slab = chain_slab + (size_t)warp_global * R * 4 * 32 + lane;
    ln.pre = bc_prefix + (size_t)warp * R * 2 * 32 + lane;
    for (;;) {
        uint32_t u = 0;
        if (lane == 0) u = atomicAdd(group_counter, 1u);
        u = __shfl_sync(0xffffffffu, u, 0);
        if (u >= total_units) break;
        // ---- task set-up (one task per lane) ----
        const uint32_t gi = u / (W * Z), rem = u - gi * (W * Z);
        const uint32_t z = rem / W, w = rem - z * W;
        const uint32_t rank = gi * 32u + lane;
        uint32_t slot = 0;
        ln.m = 0;
        ln.list = sorted;
        ln.points = z == 0 ? points0 : points1;
        if (rank < window_tasks[w]) {
            slot = order[(size_t)w * cap + rank];
            const uint2 task = tasks[(size_t)w * cap + slot];
            const uint32_t bucket = w * NB + task.x;
            const uint32_t first = task.y * T;
            ln.m = min(T, counts[bucket] - first);
            ln.list = sorted + (size_t)w * n + offsets[bucket] + first;
        }
        const uint32_t m = ln.m;
        const uint32_t mmax = __reduce_max_sync(0xffffffffu, m);
        const uint32_t steps = (mmax + R - 1) / R;

        Fq run = fp_one<FQ>();
#pragma unroll 1
        for (uint32_t s = 0; s < steps; ++s) {
            // 1 / (product of this step's denominators); nothing to invert before step 0 and for lanes whose list has ended
            Fq inv_all = run;
            if (s > 0) inv_all = fp_inv_serial(run);
            run = fp_one<FQ>();
            const bool down = (s & 1u) != 0;          // visit the chains in the opposite order of the previous step
            BcOperands nx;
            ln.request(s, down ? R - 1 : 0, nx);
#pragma unroll 1
            for (uint32_t k = 0; k < R; ++k) {
                const uint32_t r = down ? R - 1 - k : k;
                const BcOperands cu = nx;
                if (k + 1 < R) ln.request(s, down ? r - 1 : r + 1, nx);
                const uint32_t idx = s * R + r;
                if (idx >= m) continue;
                G1Affine b;
                b.x = cu.bx; b.y = cu.by;
                if (cu.eb >> 31) b.y = fp_neg(b.y);
                G1Affine o = b;                       // step 0: the chain starts with its first item
                if (s > 0) {
                    G1Affine a;
                    a.x = cu.ax; a.y = cu.ay;
                    Fq d;
                    const int kind = ba_classify(a, b, d);
                    if (kind <= 1) {
                        const Fq inv = fp_mul(inv_all, ln.pre_load(r));
                        inv_all = fp_mul(inv_all, d);
                        Fq num;
                        if (kind == 0) num = fp_sub(b.y, a.y);
                        else {
                            const Fq xx = fq_mul_call(a.x, a.x);
                            num = fp_add(fp_dbl(xx), xx);
                        }
                        const Fq lam = fp_mul(num, inv);
                        o.x = fp_sub(fp_sub(fp_mul(lam, lam), a.x), b.x);
                        o.y = fp_sub(fp_mul(lam, fp_sub(a.x, o.x)), a.y);
                    } else if (kind == 3) o = a;
                    else if (kind == 4) { o.x = fp_zero<FQ>(); o.y = fp_zero<FQ>(); }
                }
                ln.slab_store(r, o);
                // forward pass of the next step for this chain: denominator of (o + next item)
                if (idx + R < m) {
                    Fq d = fp_sub(cu.xn, o.x);
                    if (fp_is_zero(o.x) || fp_is_zero(cu.xn) || fp_is_zero(d)) {   // rare: identity operand, equal or opposite points
                        const G1Affine nb = bc_item(ln.points, cu.en);
                        if (ba_classify(o, nb, d) > 1) continue;
                    }
                    ln.pre_store(r, run);
                    run = fp_mul(run, d);
                }
            }
        }
        // tail: fold the chain sums (XYZZ mixed additions) and emit the task result
        if (m != 0) {
            G1Xyzz acc = xyzz_identity();
            const uint32_t live = min(m, (uint32_t)R);
#pragma unroll 1
            for (uint32_t r = 0; r < live; ++r) {
                G1Affine c;
                c.x = ln.slab_load(r, 0);
                c.y = ln.slab_load(r, 1);
                if (g1_affine_is_identity(c)) continue;
                ba_xyzz_madd(acc, c.x, c.y);
            }
            xyzz_store(task_out + (size_t)z * W * cap * 128, (size_t)w * cap + slot, acc);
        }
        __syncwarp();
    }
}

}  // namespace snarkv
