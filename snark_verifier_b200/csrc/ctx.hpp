// ctx.hpp — library-internal context: stream, grow-only device workspace, stage timers, error string.
// The public face of this struct is the opaque `snarkv_ctx` of include/snarkv_cuda.h (it stands where the reference
// has the global `LOADER: NativeLoader`, loader/native.rs:11-15).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/snarkv_cuda.h"

#define SNARKV_MAX_DEVICES 16   // devices of one multi-device context (one box: 8)

// global (not in the namespace): translation units built for another curve rename the namespace (msm_pasta.cu) but share snarkv_ctx
struct snarkv_stage_record {
    const char* name;
    cudaEvent_t start, stop;
    int launches;
};

namespace snarkv {

using StageRecord = ::snarkv_stage_record;

enum WsSlot : int {
    WS_POINTS_MONT = 0,   // n x 64 B   Montgomery copy of CANONICAL points
    WS_COUNTS,            // W x NB u32
    WS_OFFSETS,           // W x NB u32
    WS_CURSOR,            // W x NB u32
    WS_SORTED,            // W x n  u32
    WS_DIGITS,            // W x n  u32 (window-major signed digits)
    WS_BUCKETS,           // W x NB x 128 B
    WS_SEGPART,           // W x J  x 128 B
    WS_WINSUM,            // W x 128 B
    WS_TASK_BASE,         // W x NB u32
    WS_WINDOW_TASKS,      // W u32
    WS_TASKS,             // W x cap uint2
    WS_TASK_OUT,          // W x cap x 128 B
    WS_BIG,               // 1 + W x NB u32
    WS_ORDER,             // W x cap u32
    WS_STATUS,            // 4 B
    WS_OUT,               // small outputs (affine 64 + jacobian 96)
    WS_IO_A,              // staging for host-buffer entry points
    WS_IO_B,
    WS_IO_C,
    WS_IO_D,
    WS_IO_E,
    WS_IO_F,
    WS_BATCH_TERMS,       // batch-MSM per-term XYZZ
    WS_PAIR_A,
    WS_PAIR_B,
    WS_BA_REGION_A,       // batched-affine accumulation: level-1/3/.. scratch, B x W x ((n + cap) / 2 + 2) x 64 B
    WS_BA_REGION_B,       // level-2/4/.. scratch, half of that
    WS_BA_PREFIX,         // per resident block: K x 128 x 32 B prefix products
    WS_BA_COUNTER,        // [group counter | self-check mismatch record x 4]
    WS_TASK_OUT_CHECK,    // second task-result array (accumulate mode 3)
    WS_FR_REGS,           // fr_program register file: n_regs x m x 32 B
    WS_FR_PROG,           // fr_program: instructions | consts | out_regs
    WS_SORT_PART,         // two-level sort: [partition sizes | offsets | run cursors], 3 x W x P u32
    WS_SORT_REC_IDX,      // two-level sort: W x n u32 term references grouped by coarse partition
    WS_SORT_REC_LO,       // two-level sort: W x n u16 low digit bits of the same records
    WS_BC_SLAB,           // chained batched-affine accumulation: per resident warp R x 2 KB of running sums
    WS_SLOTS
};

}  // namespace snarkv

struct snarkv_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;   // second stream for host->device copies that overlap kernels
    cudaEvent_t copy_done[8] = {};        // one per host chunk (<= 6) + one fence
    std::string err;
    int window_bits = 0;
    int glv_mode = 0;       // 0 = GLV for n < 2^22 (default), 1 = always, 2 = never
    int pairing_mode = 0;   // 0 = choose from N, 1 = one thread per check, 2 = first cooperative kernels (block or warp per check by N), 3 = block (160 threads), 4 = warp, 5 = latency kernel (pairing_fast.cu)
    int sort_blocks_per_sm = 0;   // SNARKV_SORT_BLOCKS: resident k_partition blocks per SM (0 = default)
    int sort_tile = 8192;   // SNARKV_SORT_TILE: digits per k_partition work item (4096 | 8192); measured at 2^24: 1.30 vs 1.20 ms (profiles/r02_sort_tile_probe.txt)
    int host_chunk_min_log_n = 20, host_chunks_small = 3;   // SNARKV_HOST_CHUNK_MIN / SNARKV_HOST_CHUNKS_SMALL: chunk pipeline for inputs of 2^min .. 2^22 terms
    int host_chunks = 7, host_chunk_ratio_pct = 160;   // host entry pipeline: term-chunks of geometrically growing size (SNARKV_HOST_CHUNKS <= 7, SNARKV_HOST_RATIO in percent)
    int accumulate_mode = 0;   // 0 = choose from the bucket load, 1 = XYZZ, 2 = batched affine (tree), 3 = XYZZ + tree + task-level self-check, 4 = chained batched affine
    int ba_blocks_per_sm = 0;  // occupancy of k_bucket_accumulate_affine (queried once)
    int ba_k = 128, ba_pairs_min = 12, ba_q = 4, ba_min_load = 96;   // batched-affine tuning (developer knobs: SNARKV_BA_K, _PAIRS_MIN, _Q, _MIN_LOAD)
    int bc_r = 16, bc_blocks_per_sm = 0, bc_blocks_r = 0, bc_auto = 0, bc_min_load = 32;   // chained batched-affine kernel (SNARKV_BC_R in {8, 12, 16}, SNARKV_BC_AUTO, SNARKV_BC_MIN_LOAD)
    uint64_t launches = 0;
    int sm_count = 148;

    void* ws[snarkv::WS_SLOTS] = {};
    size_t ws_bytes[snarkv::WS_SLOTS] = {};

    bool profiling = false;
    std::vector<snarkv_stage_record> stages;       // records of the call in flight / last call
    std::vector<cudaEvent_t> event_pool;
    size_t event_next = 0;

    // IPA deciding key (pcs/ipa/decider.rs:3-16): the committing key g as resident Pallas points (capi_pasta.cu)
    void* d_ipa_g = nullptr;
    size_t ipa_n = 0;

    // Child context: its own stream and workspace slots, so that a second, independent MSM can run concurrently with one on this
    // context (the lhs / rhs MSMs of the fused PLONK batch, plonk_batch.cu).  Created on first use by ctx_aux(), tuning knobs copied.
    snarkv_ctx* aux = nullptr;
    cudaEvent_t fork_ev = nullptr, join_ev = nullptr;
    int overlap_msms = 1;   // SNARKV_OVERLAP_MSMS: 0 = run the two MSMs of the fused PLONK batch one after the other

    // KZG deciding key (pairing.cu)
    bool has_key = false;
    void* d_key_coeffs = nullptr;   // 2 x NUM_COEFFS x 3 x Fq2 line coefficients (Montgomery)
    int key_num_coeffs = 0;
    void* d_key_tables = nullptr;   // merged two-pair line constants per Miller step (pairing_fast.cu): 3 x NUM_COEFFS x 32 Fq words
    int pf_blocks_per_sm = 0;       // occupancy of k_kzg_decide_fast (queried once)
    int pf_max_per_sm = 2;          // SNARKV_PAIRING_FAST_MAX: automatic mode gives a check a whole block up to this many checks per SM
    uint8_t key_g1[64] = {};

    int fail(int code, const char* what, cudaError_t ce = cudaSuccess) {
        char buf[512];
        if (ce != cudaSuccess) snprintf(buf, sizeof buf, "%s: %s (%s)", what, cudaGetErrorName(ce), cudaGetErrorString(ce));
        else snprintf(buf, sizeof buf, "%s", what);
        err = buf;
        return code;
    }
    // grow-only workspace slot; returns nullptr (and sets err) on allocation failure
    void* wsget(int slot, size_t bytes) {
        if (bytes == 0) bytes = 16;
        if (ws_bytes[slot] >= bytes) return ws[slot];
        if (ws[slot]) {
            cudaStreamSynchronize(stream);
            cudaFree(ws[slot]);
            ws[slot] = nullptr;
            ws_bytes[slot] = 0;
        }
        size_t want = bytes + bytes / 8;  // slack so a sweep of growing sizes does not reallocate every call
        cudaError_t ce = cudaMalloc(&ws[slot], want);
        if (ce != cudaSuccess) {
            ce = cudaMalloc(&ws[slot], bytes);
            want = bytes;
        }
        if (ce != cudaSuccess) {
            fail(SNARKV_ERR_CUDA, "cudaMalloc(workspace)", ce);
            ws[slot] = nullptr;
            return nullptr;
        }
        ws_bytes[slot] = want;
        return ws[slot];
    }
    cudaEvent_t next_event() {
        if (event_next == event_pool.size()) {
            cudaEvent_t e;
            cudaEventCreate(&e);
            event_pool.push_back(e);
        }
        return event_pool[event_next++];
    }
    void profile_begin_call() {
        stages.clear();
        event_next = 0;
    }
};

namespace snarkv {

// RAII stage bracket: records start/stop events on the context's stream when profiling is on.
struct Stage {
    snarkv_ctx* c;
    int idx = -1;
    Stage(snarkv_ctx* ctx, const char* name) : c(ctx) {
        if (!c->profiling) return;
        StageRecord r{name, c->next_event(), c->next_event(), 0};
        cudaEventRecord(r.start, c->stream);
        c->stages.push_back(r);
        idx = (int)c->stages.size() - 1;
    }
    void launched(int k = 1) {
        c->launches += (uint64_t)k;
        if (idx >= 0) c->stages[idx].launches += k;
    }
    ~Stage() {
        if (idx >= 0) cudaEventRecord(c->stages[idx].stop, c->stream);
    }
};

#define SNARKV_CUDA_TRY(ctx, expr)                                              \
    do {                                                                         \
        cudaError_t _ce = (expr);                                                \
        if (_ce != cudaSuccess) return (ctx)->fail(SNARKV_ERR_CUDA, #expr, _ce); \
    } while (0)
#define SNARKV_LAUNCH_CHECK(ctx, name)                                             \
    do {                                                                            \
        cudaError_t _ce = cudaGetLastError();                                       \
        if (_ce != cudaSuccess) return (ctx)->fail(SNARKV_ERR_CUDA, name, _ce);     \
    } while (0)

// internal entry points shared between translation units
int msm_run_device(snarkv_ctx* ctx, const void* d_scalars, const void* d_points, size_t n, int scalar_format, int point_format,
                   int out_format, int flags, void* d_out_affine, void* d_out_jacobian, void* d_status);
int msm_run_device_pair(snarkv_ctx* ctx, const void* d_scalars, const void* d_points0, const void* d_points1, size_t n, int scalar_format,
                        int point_format, int out_format, int flags, void* d_out_affine, void* d_status);
int msm_run_host(snarkv_ctx* ctx, const uint8_t* scalars, const uint8_t* points, size_t n, int format, int flags, uint8_t* out,
                 void* d_out_jacobian, const uint8_t* d_resident = nullptr);
int msm_bases_prepare(snarkv_ctx* ctx, const void* d_in, size_t n, int format, int check, void* d_out, void* d_status);
void msm_plan_query(snarkv_ctx* ctx, size_t n, uint32_t out[4]);
int msm_fold_partials_device(snarkv_ctx* ctx, const void* d_partials, size_t k, int format, void* d_out_affine);
int msm_batch_device(snarkv_ctx* ctx, const void* d_scalars, const void* d_points, const void* d_offsets, size_t m, size_t total,
                     int format, int flags, void* d_out_affine, void* d_status);
int msm_batch_rlc_device(snarkv_ctx* ctx, const void* d_scalars, const void* d_points, const void* d_offsets, size_t m, size_t total,
                         const void* d_rho, int format, int flags, void* d_scaled, void* d_powers, void* d_out_affine, void* d_status,
                         uint64_t first_power = 0, void* d_out_jacobian = nullptr);
int msm_batch_rlc_host(snarkv_ctx* ctx, const uint8_t* scalars, const uint8_t* points, const uint64_t* offsets, size_t m, const uint8_t rho[32],
                       int format, int flags, uint64_t first_power, uint8_t* out_affine, void* d_out_jacobian);
int msm_fold_partials_peer(snarkv_ctx* ctx, const void* const* d_partials, size_t k, int format, void* d_out_affine);
int fr_batch_invert_device(snarkv_ctx* ctx, void* d_values, size_t n, int format, const void* d_coeff, void* d_scratch);
int accumulators_from_limbs_device(snarkv_ctx* ctx, const void* d_limbs, size_t m, uint32_t L, uint32_t bits, int format, void* d_lhs,
                                   void* d_rhs, void* d_valid);
int fr_program_device(snarkv_ctx* ctx, const void* d_prog, size_t n_instr, void* d_consts, size_t n_consts, const void* d_inputs,
                      size_t n_inputs, size_t m, int format, uint32_t n_regs, const void* d_out_regs, size_t n_out, void* d_outputs);
int fr_mul_vec_device(snarkv_ctx* ctx, const void* d_a, const void* d_b, size_t n, int format, void* d_out);
int fr_from_mont_device(snarkv_ctx* ctx, void* d_v, size_t n);
int evm_transcript_device(snarkv_ctx* ctx, const void* d_streams, size_t stream_len, const void* d_seg_end, size_t k, size_t m, int format,
                          void* d_out);
int poseidon_transcript_device(snarkv_ctx* ctx, const void* d_elements, size_t stream_len, const void* d_seg_end, size_t k, size_t m, int format,
                               void* d_out);
int plonk_poseidon_expand_device(snarkv_ctx* ctx, const void* d_proofs, uint32_t n_items, const void* d_item_off, const void* d_item_pt, uint32_t stream_len,
                                 uint32_t n_points, size_t m, void* d_elements, void* d_points, void* d_status, uint32_t in_stride, uint32_t in_base);
int fr_powers_device(snarkv_ctx* ctx, const void* d_r, int format, size_t n, void* d_out_mont, uint64_t first = 0);
int field_op_device(snarkv_ctx* ctx, int field, int op, const void* d_a, const void* d_b, size_t n, void* d_out);
int synth_scalars_device(snarkv_ctx* ctx, uint64_t seed, uint64_t start, size_t n, int format, void* d_out);
int synth_points_device(snarkv_ctx* ctx, uint64_t seed, uint64_t start, size_t n, int format, void* d_out);
int kzg_set_key(snarkv_ctx* ctx, const uint8_t g1[64], const uint8_t g2[128], const uint8_t s_g2[128]);
int kzg_decide_device(snarkv_ctx* ctx, const void* d_lhs, const void* d_rhs, size_t N, int format, void* d_accept, void* d_gt);
int kzg_decide_coop_device(snarkv_ctx* ctx, const void* d_lhs, const void* d_rhs, size_t N, int format, void* d_accept, void* d_gt);
int kzg_decide_fast_device(snarkv_ctx* ctx, const void* d_lhs, const void* d_rhs, size_t N, int format, void* d_accept, void* d_gt);
int kzg_build_pair_tables(snarkv_ctx* ctx);
snarkv_ctx* ctx_aux(snarkv_ctx* ctx);   // capi.cu: the child context (nullptr + ctx->err on failure)
void kzg_free_key(snarkv_ctx* ctx);

}  // namespace snarkv
