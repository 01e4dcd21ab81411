// plonk_batch.cu — device-resident batch verification of m PLONK proofs of ONE protocol from their absorbed byte streams
// (SURVEY.md §8 config 3 "one fused G1 MSM + one batched multi-pairing", rows a11 / a12 / f2 / f3 in one call).
//
// What the reference does per proof on the host (paths relative to snark-verifier/src) and where it runs here:
//   PlonkProof::read                     verifier/plonk/proof.rs:52-169     transcript kernel (transcript.cu) + k_plonk_rows / k_plonk_points:
//                                                                           the proof is PARSED on the device — 32-byte big-endian words ->
//                                                                           little-endian scalars (`read_scalar`: reject >= r) and points
//   PlonkSuccinctVerifier::verify        verifier/plonk.rs:57-93            ONE straight-line Fr program per protocol (fr_program.cu), compiled
//     (common polynomials, evaluations,                                     on the host by snark_verifier_b200.plonk.compile_plonk_verifier
//      commitments, queries, Gwc19 / Bdfg21::verify)
//   Msm::evaluate x 2 per proof          util/msm.rs:81-98                  sum_j rho^j MSM_j as ONE Pippenger pass per side (msm.cu), every
//                                                                           proof point validated (`from_xy`: SNARKV_CHECK_INPUTS)
//   decide_all                           pcs/kzg/decider.rs:84-93, 146-185  one pairing check on the fused accumulator (pairing*.cu)
// Nothing but the streams goes up and nothing but the 128-byte accumulator and the verdict comes back.
#include "ctx.hpp"
#include "fp.cuh"

struct snarkv_plonk_plan {
    int device = 0;
    uint32_t transcript = 0, n_pre = 0, n_items = 0, n_points = 0;   // Poseidon: caller-supplied leading elements, proof items, points per proof
    int32_t *d_item_off = nullptr, *d_item_pt = nullptr;
    uint32_t stream_words = 0, n_challenges = 0, n_instr = 0, n_regs = 0, n_consts = 0, n_inputs = 0, n_out = 0, n_lhs = 0, n_rhs = 0;
    // device tables (one allocation)
    uint8_t* d_tables = nullptr;
    uint32_t* d_seg_end = nullptr;
    uint8_t *d_prog = nullptr, *d_consts = nullptr, *d_consts_work = nullptr, *d_const_points = nullptr;   // d_consts: pristine canonical copy
    uint32_t* d_out_regs = nullptr;
    int32_t *d_row_src = nullptr, *d_lhs_src = nullptr, *d_rhs_src = nullptr;
    uint8_t* d_row_check = nullptr;
    // per-call buffers, grown on demand
    uint8_t* d_work = nullptr;
    size_t work_bytes = 0;
};

namespace snarkv {

// 32-byte big-endian word -> little-endian limbs
__device__ __forceinline__ Fr load_be_word(const uint8_t* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    const uint4 hi = __ldg(q), lo = __ldg(q + 1);   // hi holds the most significant 16 bytes
    Fr r;
    r.v[0] = __byte_perm(lo.w, 0, 0x0123); r.v[1] = __byte_perm(lo.z, 0, 0x0123); r.v[2] = __byte_perm(lo.y, 0, 0x0123); r.v[3] = __byte_perm(lo.x, 0, 0x0123);
    r.v[4] = __byte_perm(hi.w, 0, 0x0123); r.v[5] = __byte_perm(hi.z, 0, 0x0123); r.v[6] = __byte_perm(hi.y, 0, 0x0123); r.v[7] = __byte_perm(hi.x, 0, 0x0123);
    return r;
}

// program input rows: rows[j][i] = the proof's word row_src[i] (little-endian) or its challenge -(row_src[i] + 1)
__global__ void __launch_bounds__(256) k_plonk_rows(const uint8_t* __restrict__ streams, uint32_t stream_words, const uint8_t* __restrict__ challenges,
                                                    uint32_t n_challenges, const int32_t* __restrict__ row_src, const uint8_t* __restrict__ row_check,
                                                    uint32_t n_inputs, size_t m, int little_endian, uint8_t* __restrict__ rows,
                                                    int* __restrict__ status) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m * n_inputs) return;
    const size_t j = t / n_inputs;
    const uint32_t i = (uint32_t)(t - j * n_inputs);
    const int32_t src = row_src[i];
    Fr v;
    if (src >= 0) {
        const uint8_t* w = streams + (j * stream_words + (uint32_t)src) * 32;
        v = little_endian ? fp_load<FR>(w) : load_be_word(w);     // Poseidon element stream: already little-endian scalars
        if (row_check[i] && !fp_is_canonical(v)) atomicCAS(status, 0, SNARKV_ERR_BAD_SCALAR);   // read_scalar: "Invalid scalar encoding in proof"
    } else {
        v = fp_load<FR>(challenges + (j * n_challenges + (uint32_t)(-(src + 1))) * 32);
    }
    fp_store<FR>(rows + t * 32, v);
}

// bases of one MSM side: pts[j][k] = constant point -(src[k] + 1) or the proof's point at words src[k], src[k] + 1; scal[j][k] = outputs[j][first + k]
__global__ void __launch_bounds__(256) k_plonk_side(const uint8_t* __restrict__ streams, uint32_t stream_words, const uint8_t* __restrict__ const_points,
                                                    const int32_t* __restrict__ src, uint32_t n_slots, const uint8_t* __restrict__ outputs,
                                                    uint32_t n_out, uint32_t first, size_t m, const uint8_t* __restrict__ parsed_points,
                                                    uint32_t n_points, uint8_t* __restrict__ pts, uint8_t* __restrict__ scal) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m * n_slots) return;
    const size_t j = t / n_slots;
    const uint32_t k = (uint32_t)(t - j * n_slots);
    const int32_t s = src[k];
    Fr x, y;
    if (s >= 0 && parsed_points) {                                 // Poseidon transcript: the proof's points were decompressed into an array
        const uint8_t* w = parsed_points + (j * n_points + (uint32_t)s) * 64;
        x = fp_load<FR>(w);
        y = fp_load<FR>(w + 32);
    } else if (s >= 0) {
        const uint8_t* w = streams + (j * stream_words + (uint32_t)s) * 32;
        x = load_be_word(w);
        y = load_be_word(w + 32);
    } else {
        const uint8_t* c = const_points + (size_t)(-(s + 1)) * 64;
        x = fp_load<FR>(c);
        y = fp_load<FR>(c + 32);
    }
    fp_store<FR>(pts + t * 64, x);
    fp_store<FR>(pts + t * 64 + 32, y);
    fp_store<FR>(scal + t * 32, fp_load<FR>(outputs + (j * n_out + first + k) * 32));
}

__global__ void k_plonk_offsets(uint64_t* __restrict__ off, size_t m, uint32_t per) {
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j <= m) off[j] = (uint64_t)j * per;
}

}  // namespace snarkv

using namespace snarkv;

#define BCTX_GUARD(ctx)                                                                  \
    do {                                                                                 \
        if (!(ctx)) return SNARKV_ERR_USAGE;                                             \
        (ctx)->err.clear();                                                              \
        cudaError_t _g = cudaSetDevice((ctx)->device);                                   \
        if (_g != cudaSuccess) return (ctx)->fail(SNARKV_ERR_CUDA, "cudaSetDevice", _g); \
    } while (0)

static size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

extern "C" {

int snarkv_plonk_plan_create(snarkv_ctx* ctx, const snarkv_plonk_plan_desc* d, snarkv_plonk_plan** out) {
    BCTX_GUARD(ctx);
    if (!d || !out) return ctx->fail(SNARKV_ERR_USAGE, "snarkv_plonk_plan_create: bad argument");
    *out = nullptr;
    if (d->transcript > 1) return ctx->fail(SNARKV_ERR_USAGE, "snarkv_plonk_plan_create: transcript 0 (Keccak EvmTranscript) or 1 (Poseidon)");
    const bool pos = d->transcript == 1;
    if (pos && (!d->item_off || !d->item_pt || d->n_items == 0 || d->n_pre + d->n_items == 0))
        return ctx->fail(SNARKV_ERR_USAGE, "snarkv_plonk_plan_create: the Poseidon transcript needs the proof item table");
    if (pos) {
        for (uint32_t i = 0; i < d->n_items; ++i) {
            const int32_t o = d->item_off[i];
            const uint32_t off = o >= 0 ? (uint32_t)o : (uint32_t)(-(o + 1)) + 1;
            if (off >= d->stream_words || (o < 0 && (d->item_pt[i] < 0 || (uint32_t)d->item_pt[i] >= d->n_points)))
                return ctx->fail(SNARKV_ERR_USAGE, "snarkv_plonk_plan_create: proof item out of range");
        }
        if (d->n_pre > d->stream_words) return ctx->fail(SNARKV_ERR_USAGE, "snarkv_plonk_plan_create: n_pre exceeds the stream");
    }
    if (!d->seg_end || d->n_challenges == 0 || !d->program || d->n_instr == 0 || d->n_regs == 0 || !d->out_regs || d->n_out == 0 || !d->row_src ||
        !d->row_check || d->n_inputs == 0 || d->n_lhs + d->n_rhs != d->n_out || d->n_lhs == 0 || d->n_rhs == 0 || !d->lhs_src || !d->rhs_src ||
        (d->n_consts && !d->consts) || (d->n_const_points && !d->const_points) || d->stream_words == 0)
        return ctx->fail(SNARKV_ERR_USAGE, "snarkv_plonk_plan_create: incomplete description");
    uint32_t prev = 0;
    for (uint32_t i = 0; i < d->n_challenges; ++i) {
        if (d->seg_end[i] < prev || d->seg_end[i] > d->stream_words) return ctx->fail(SNARKV_ERR_USAGE, "seg_end must be non-decreasing word offsets within the stream");
        prev = d->seg_end[i];
    }
    for (uint32_t i = 0; i < d->n_inputs; ++i) {
        const int32_t s = d->row_src[i];
        if (s >= 0 ? (uint32_t)s >= d->stream_words : (uint32_t)(-(s + 1)) >= d->n_challenges) return ctx->fail(SNARKV_ERR_USAGE, "row_src out of range");
    }
    for (int side = 0; side < 2; ++side) {
        const int32_t* src = side ? d->rhs_src : d->lhs_src;
        const uint32_t n = side ? d->n_rhs : d->n_lhs;
        for (uint32_t k = 0; k < n; ++k) {
            const int32_t s = src[k];
            if (s >= 0 ? (pos ? (uint32_t)s >= d->n_points : (uint32_t)s + 1 >= d->stream_words) : (uint32_t)(-(s + 1)) >= d->n_const_points)
                return ctx->fail(SNARKV_ERR_USAGE, "point source out of range");
        }
    }
    // instruction validation as in snarkv_fr_program_eval_batch: opcodes, operand ranges, no read before write
    {
        std::vector<uint8_t> written(d->n_regs, 0);
        for (size_t i = 0; i < d->n_instr; ++i) {
            const snarkv_fr_instr& in = d->program[i];
            bool ok = in.dst < d->n_regs;
            auto src = [&](uint32_t r) { return r < d->n_regs && written[r]; };
            switch (in.op) {
                case SNARKV_FR_OP_INPUT: ok = ok && in.a < d->n_inputs; break;
                case SNARKV_FR_OP_CONST: ok = ok && in.a < d->n_consts; break;
                case SNARKV_FR_OP_ADD: case SNARKV_FR_OP_SUB: case SNARKV_FR_OP_MUL: case SNARKV_FR_OP_KEEPZ: ok = ok && src(in.a) && src(in.b); break;
                case SNARKV_FR_OP_NEG: case SNARKV_FR_OP_INV: case SNARKV_FR_OP_NZ: ok = ok && src(in.a); break;
                default: ok = false;
            }
            if (!ok) return ctx->fail(SNARKV_ERR_USAGE, "snarkv_plonk_plan_create: invalid program instruction");
            written[in.dst] = 1;
        }
        for (uint32_t k = 0; k < d->n_out; ++k)
            if (d->out_regs[k] >= d->n_regs || !written[d->out_regs[k]]) return ctx->fail(SNARKV_ERR_USAGE, "snarkv_plonk_plan_create: output register never written");
    }
    snarkv_plonk_plan* p = new snarkv_plonk_plan();
    p->device = ctx->device;
    p->transcript = d->transcript; p->n_pre = pos ? d->n_pre : 0; p->n_items = pos ? d->n_items : 0; p->n_points = pos ? d->n_points : 0;
    p->stream_words = d->stream_words; p->n_challenges = d->n_challenges; p->n_instr = (uint32_t)d->n_instr; p->n_regs = d->n_regs;
    p->n_consts = (uint32_t)d->n_consts; p->n_inputs = d->n_inputs; p->n_out = d->n_out; p->n_lhs = d->n_lhs; p->n_rhs = d->n_rhs;
    const size_t b_seg = al256(d->n_challenges * 4), b_prog = al256(d->n_instr * sizeof(snarkv_fr_instr)), b_consts = al256(d->n_consts * 32 + 32),
                 b_outr = al256(d->n_out * 4), b_rows = al256(d->n_inputs * 4), b_chk = al256(d->n_inputs), b_l = al256(d->n_lhs * 4), b_r = al256(d->n_rhs * 4),
                 b_cp = al256(d->n_const_points * 64 + 64), b_items = pos ? al256(d->n_items * 4) : 0;
    const size_t total = 2 * b_items + b_seg + b_prog + 2 * b_consts + b_outr + b_rows + b_chk + b_l + b_r + b_cp;
    cudaError_t ce = cudaMalloc(&p->d_tables, total);
    if (ce != cudaSuccess) { delete p; return ctx->fail(SNARKV_ERR_CUDA, "cudaMalloc(plan tables)", ce); }
    std::vector<uint8_t> host(total, 0);
    size_t o = 0;
    auto put = [&](const void* src, size_t bytes, size_t slot) { uint8_t* dev = p->d_tables + o; if (bytes) memcpy(host.data() + o, src, bytes); o += slot; return dev; };
    {
        std::vector<uint32_t> seg_bytes(d->seg_end, d->seg_end + d->n_challenges);   // the transcript kernel cuts the stream at byte offsets
        if (!pos) for (auto& e : seg_bytes) e *= 32;             // Poseidon: element offsets as they are
        p->d_seg_end = (uint32_t*)put(seg_bytes.data(), d->n_challenges * 4, b_seg);
    }
    p->d_prog = put(d->program, d->n_instr * sizeof(snarkv_fr_instr), b_prog);
    p->d_consts = put(d->consts, d->n_consts * 32, b_consts);
    p->d_consts_work = put(nullptr, 0, b_consts);
    p->d_out_regs = (uint32_t*)put(d->out_regs, d->n_out * 4, b_outr);
    p->d_row_src = (int32_t*)put(d->row_src, d->n_inputs * 4, b_rows);
    p->d_row_check = put(d->row_check, d->n_inputs, b_chk);
    p->d_lhs_src = (int32_t*)put(d->lhs_src, d->n_lhs * 4, b_l);
    p->d_rhs_src = (int32_t*)put(d->rhs_src, d->n_rhs * 4, b_r);
    p->d_const_points = put(d->const_points, d->n_const_points * 64, b_cp);
    if (pos) {
        p->d_item_off = (int32_t*)put(d->item_off, d->n_items * 4, b_items);
        p->d_item_pt = (int32_t*)put(d->item_pt, d->n_items * 4, b_items);
    }
    ce = cudaMemcpy(p->d_tables, host.data(), total, cudaMemcpyHostToDevice);
    if (ce != cudaSuccess) { cudaFree(p->d_tables); delete p; return ctx->fail(SNARKV_ERR_CUDA, "cudaMemcpy(plan tables)", ce); }
    *out = p;
    return SNARKV_OK;
}

void snarkv_plonk_plan_free(snarkv_ctx* ctx, snarkv_plonk_plan* p) {
    if (!p) return;
    if (ctx) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream); }
    if (p->d_tables) cudaFree(p->d_tables);
    if (p->d_work) cudaFree(p->d_work);
    delete p;
}

int snarkv_plonk_accumulate_batch(snarkv_ctx* ctx, snarkv_plonk_plan* p, const uint8_t* streams, size_t m, const uint8_t rho[32], int decide,
                                  uint8_t out_lhs[64], uint8_t out_rhs[64], uint8_t* accept) {
    BCTX_GUARD(ctx);
    if (!p || !streams || !rho || m == 0 || (decide && !accept)) return ctx->fail(SNARKV_ERR_USAGE, "snarkv_plonk_accumulate_batch: bad argument");
    if (p->device != ctx->device) return ctx->fail(SNARKV_ERR_USAGE, "plan lives on another device");
    if (decide && !ctx->has_key) return ctx->fail(SNARKV_ERR_NO_KEY, "snarkv_plonk_accumulate_batch: no deciding key installed");
    ctx->profile_begin_call();
    const size_t sw = p->stream_words, nl = p->n_lhs, nr = p->n_rhs;
    // per-call buffers: streams | challenges | rows | outputs | lhs pts | lhs scal | rhs pts | rhs scal | offsets x 2 | scaled x 2 | powers | rho | out
    const size_t b_st = al256(m * sw * 32), b_ch = al256(m * p->n_challenges * 32), b_rows = al256(m * p->n_inputs * 32 + 32), b_out = al256(m * p->n_out * 32),
                 b_lp = al256(m * nl * 64), b_ls = al256(m * nl * 32), b_rp = al256(m * nr * 64), b_rs = al256(m * nr * 32), b_off = al256((m + 1) * 8),
                 b_pow = 2 * al256(m * 32), b_misc = 1024;   // two copies of the powers of rho: the two MSMs may run concurrently
    const bool pos = p->transcript == 1;
    const size_t in_words = pos ? (size_t)p->n_pre + p->n_items : sw;         // what the caller uploads per proof
    const size_t b_in = pos ? al256(m * in_words * 32) : 0, b_pts = pos ? al256(m * (size_t)p->n_points * 64 + 64) : 0;
    const size_t need = b_in + b_pts + b_st + b_ch + b_rows + b_out + b_lp + 2 * b_ls + b_rp + 2 * b_rs + 2 * b_off + b_pow + b_misc;
    if (need > p->work_bytes) {
        if (p->d_work) { cudaStreamSynchronize(ctx->stream); cudaFree(p->d_work); p->d_work = nullptr; p->work_bytes = 0; }
        SNARKV_CUDA_TRY(ctx, cudaMalloc(&p->d_work, need + need / 4));
        p->work_bytes = need + need / 4;
    }
    uint8_t* w = p->d_work;
    uint8_t* d_in = w; w += b_in;
    uint8_t* d_pts = pos ? w : nullptr; w += b_pts;
    uint8_t* d_st = w; w += b_st;
    uint8_t* d_ch = w; w += b_ch;
    uint8_t* d_rows = w; w += b_rows;
    uint8_t* d_out = w; w += b_out;
    uint8_t* d_lp = w; w += b_lp;
    uint8_t* d_ls = w; w += b_ls;
    uint8_t* d_lsc = w; w += b_ls;
    uint8_t* d_rp = w; w += b_rp;
    uint8_t* d_rs = w; w += b_rs;
    uint8_t* d_rsc = w; w += b_rs;
    uint64_t* d_loff = (uint64_t*)w; w += b_off;
    uint64_t* d_roff = (uint64_t*)w; w += b_off;
    uint8_t* d_pow = w; w += b_pow;
    uint8_t* d_misc = w;                                   // [rho 32 | lhs 64 | rhs 64 | accept 1 @192 | status words @256..]
    cudaStream_t st = ctx->stream;
    SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(pos ? d_in : d_st, streams, m * in_words * 32, cudaMemcpyHostToDevice, st));
    SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(d_misc, rho, 32, cudaMemcpyHostToDevice, st));
    SNARKV_CUDA_TRY(ctx, cudaMemsetAsync(d_misc + 256, 0, 64, st));
    int* d_status = (int*)(d_misc + 256);                  // [0]: rows / parsing; [2..3]: lhs rlc + msm; [4..5]: rhs rlc + msm
    int rc;
    if (pos) {
        // element stream = [caller's leading elements | what every proof item absorbs]; compressed points are decompressed + validated here
        if (p->n_pre)
            SNARKV_CUDA_TRY(ctx, cudaMemcpy2DAsync(d_st, sw * 32, d_in, in_words * 32, (size_t)p->n_pre * 32, m, cudaMemcpyDeviceToDevice, st));
        rc = plonk_poseidon_expand_device(ctx, d_in, p->n_items, p->d_item_off, p->d_item_pt, (uint32_t)sw, p->n_points, m, d_st, d_pts, d_status,
                                          (uint32_t)in_words, p->n_pre);
        if (rc) return rc;
        rc = poseidon_transcript_device(ctx, d_st, sw, p->d_seg_end, p->n_challenges, m, SNARKV_CANONICAL, d_ch);      // seg_end: ELEMENT offsets
    } else {
        rc = evm_transcript_device(ctx, d_st, sw * 32, p->d_seg_end, p->n_challenges, m, SNARKV_CANONICAL, d_ch);      // seg_end: BYTE offsets
    }
    if (rc) return rc;
    {
        Stage sg(ctx, "plonk_parse_rows");
        const size_t n = m * p->n_inputs;
        k_plonk_rows<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_st, (uint32_t)sw, d_ch, p->n_challenges, p->d_row_src, p->d_row_check, p->n_inputs, m,
                                                                 pos ? 1 : 0, d_rows, d_status);
        SNARKV_LAUNCH_CHECK(ctx, "k_plonk_rows");
        sg.launched();
    }
    // fr_program_device converts canonical constants to Montgomery form IN PLACE: hand it a fresh copy every call
    if (p->n_consts) SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(p->d_consts_work, p->d_consts, (size_t)p->n_consts * 32, cudaMemcpyDeviceToDevice, st));
    rc = fr_program_device(ctx, p->d_prog, p->n_instr, p->d_consts_work, p->n_consts, d_rows, p->n_inputs, m, SNARKV_CANONICAL, p->n_regs, p->d_out_regs, p->n_out, d_out);
    if (rc) return rc;
    {
        Stage sg(ctx, "plonk_parse_points");
        k_plonk_side<<<(unsigned)((m * nl + 255) / 256), 256, 0, st>>>(d_st, (uint32_t)sw, p->d_const_points, p->d_lhs_src, (uint32_t)nl, d_out, p->n_out, 0, m, d_pts, p->n_points, d_lp, d_ls);
        SNARKV_LAUNCH_CHECK(ctx, "k_plonk_side");
        k_plonk_side<<<(unsigned)((m * nr + 255) / 256), 256, 0, st>>>(d_st, (uint32_t)sw, p->d_const_points, p->d_rhs_src, (uint32_t)nr, d_out, p->n_out, (uint32_t)nl, m, d_pts, p->n_points, d_rp, d_rs);
        SNARKV_LAUNCH_CHECK(ctx, "k_plonk_side");
        k_plonk_offsets<<<(unsigned)((m + 256) / 256), 256, 0, st>>>(d_loff, m, (uint32_t)nl);
        k_plonk_offsets<<<(unsigned)((m + 256) / 256), 256, 0, st>>>(d_roff, m, (uint32_t)nr);
        SNARKV_LAUNCH_CHECK(ctx, "k_plonk_offsets");
        sg.launched(4);
    }
    // The two MSMs are independent and, at a few 10^4 terms each, latency-bound chains of small kernels (scan, bucket reduction,
    // Horner tail ~0.9 ms each): the rhs one runs on the child context's stream next to the lhs one (measured: 4096 proofs
    // 3.67 -> see profiles/), joined before the pairing.  With stage profiling on they run one after the other on this stream so
    // that the CUDA-event brackets stay meaningful.
    snarkv_ctx* ax = (ctx->overlap_msms && !ctx->profiling) ? ctx_aux(ctx) : nullptr;
    if (ax) {
        SNARKV_CUDA_TRY(ctx, cudaEventRecord(ctx->fork_ev, st));
        SNARKV_CUDA_TRY(ctx, cudaStreamWaitEvent(ax->stream, ctx->fork_ev, 0));
        ax->err.clear();
        rc = msm_batch_rlc_device(ax, d_rs, d_rp, d_roff, m, m * nr, d_misc, SNARKV_CANONICAL, SNARKV_CHECK_INPUTS, d_rsc, d_pow + al256(m * 32), d_misc + 96,
                                  d_status + 4);
        ctx->launches += ax->launches;
        ax->launches = 0;
        if (rc) { cudaStreamSynchronize(ax->stream); return ctx->fail(rc, ax->err.c_str()); }
        SNARKV_CUDA_TRY(ctx, cudaEventRecord(ctx->join_ev, ax->stream));
    }
    rc = msm_batch_rlc_device(ctx, d_ls, d_lp, d_loff, m, m * nl, d_misc, SNARKV_CANONICAL, SNARKV_CHECK_INPUTS, d_lsc, d_pow, d_misc + 32, d_status + 2);
    if (rc) { if (ax) cudaStreamSynchronize(ax->stream); return rc; }
    if (ax) {
        SNARKV_CUDA_TRY(ctx, cudaStreamWaitEvent(st, ctx->join_ev, 0));
    } else {
        rc = msm_batch_rlc_device(ctx, d_rs, d_rp, d_roff, m, m * nr, d_misc, SNARKV_CANONICAL, SNARKV_CHECK_INPUTS, d_rsc, d_pow + al256(m * 32), d_misc + 96,
                                  d_status + 4);
        if (rc) return rc;
    }
    if (decide) {
        rc = kzg_decide_device(ctx, d_misc + 32, d_misc + 96, 1, SNARKV_CANONICAL, d_misc + 192, nullptr);
        if (rc) return rc;
    }
    uint8_t host[256 + 64];
    SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(host, d_misc, sizeof host, cudaMemcpyDeviceToHost, st));
    SNARKV_CUDA_TRY(ctx, cudaStreamSynchronize(st));
    const int* hs = (const int*)(host + 256);
    if (hs[0]) return ctx->fail(hs[0], hs[0] == SNARKV_ERR_BAD_POINT ? "Invalid elliptic curve point encoding in proof" : "Invalid scalar encoding in proof");
    for (int k = 2; k < 6; ++k)
        if (hs[k]) return ctx->fail(hs[k], hs[k] == SNARKV_ERR_BAD_SCALAR ? "scalar is not a canonical Fr" : "Invalid elliptic curve point encoding in proof");
    if (out_lhs) memcpy(out_lhs, host + 32, 64);
    if (out_rhs) memcpy(out_rhs, host + 96, 64);
    if (decide) accept[0] = host[192];
    return SNARKV_OK;
}

}  // extern "C"
