// poseidon.cu — batched Poseidon transcript challenges and compressed-point parsing.
//
// Replaces, for a batch of proofs that share one transcript shape (paths relative to snark-verifier/src):
//   util/hash/poseidon.rs:117-203                 Poseidon<F, L, T, RATE>::{update, squeeze} (T = 5, RATE = 4, R_F = 8, R_P = 60:
//                                                 what snark-verifier-sdk/src/halo2.rs:53-56 instantiates)
//   system/halo2/transcript/halo2.rs:201-242      native PoseidonTranscript::{common_scalar, common_ec_point, squeeze_challenge}
//   system/halo2/transcript/halo2.rs:244-274      read_scalar (32-byte little-endian repr) / read_ec_point (`C::from_bytes`, compressed)
// A proof's transcript is the stream of scalar-field ELEMENTS it absorbs (a scalar = 1 element, a point = x mod r, y mod r) cut at
// the squeeze points.  Fiat-Shamir is sequential inside one proof and independent across proofs: one thread per proof.
// The permutation is the textbook one (add round constants, x^5 on all / the first word, MDS); the reference's optimised form
// (pre-sparse / sparse matrices) computes the same function.  Constants: poseidon_consts.inc (gen_poseidon_consts.py: Grain LFSR,
// verified against the public Poseidon test vectors at generation time).
#include "ctx.hpp"
#include "g1.cuh"
#include "poseidon_consts.inc"

namespace snarkv {

__device__ const uint32_t POSEIDON_RC[(SNARKV_POSEIDON_RF + SNARKV_POSEIDON_RP) * SNARKV_POSEIDON_T][8] = SNARKV_POSEIDON_RC_INIT;
__device__ const uint32_t POSEIDON_MDS[SNARKV_POSEIDON_T * SNARKV_POSEIDON_T][8] = SNARKV_POSEIDON_MDS_INIT;

__device__ __forceinline__ Fr fr_const(const uint32_t (&l)[8]) {
    Fr r;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.v[i] = l[i];
    return r;
}
__device__ __forceinline__ Fr fr_pow5(const Fr& x) {
    const Fr x2 = fp_sqr(x);
    return fp_mul(fp_sqr(x2), x);
}

// state <- Poseidon permutation(state), Montgomery form
__device__ void poseidon_permute(Fr (&s)[SNARKV_POSEIDON_T]) {
    constexpr int T = SNARKV_POSEIDON_T, RF = SNARKV_POSEIDON_RF, RP = SNARKV_POSEIDON_RP;
#pragma unroll 1
    for (int r = 0; r < RF + RP; ++r) {
#pragma unroll
        for (int i = 0; i < T; ++i) s[i] = fp_add(s[i], fr_const(POSEIDON_RC[r * T + i]));
        if (r < RF / 2 || r >= RF / 2 + RP) {
#pragma unroll
            for (int i = 0; i < T; ++i) s[i] = fr_pow5(s[i]);
        } else {
            s[0] = fr_pow5(s[0]);
        }
        Fr n[T];
#pragma unroll
        for (int i = 0; i < T; ++i) {
            Fr acc = fp_mul(fr_const(POSEIDON_MDS[i * T]), s[0]);
#pragma unroll
            for (int j = 1; j < T; ++j) acc = fp_add(acc, fp_mul(fr_const(POSEIDON_MDS[i * T + j]), s[j]));
            n[i] = acc;
        }
#pragma unroll
        for (int i = 0; i < T; ++i) s[i] = n[i];
    }
}

// util/hash/poseidon.rs:46-80 + :159-173: add up to RATE inputs to state[1..], a padding 1 behind them, permute
__device__ __forceinline__ void poseidon_absorb(Fr (&s)[SNARKV_POSEIDON_T], const uint8_t* elems, uint32_t cnt, int format) {
    constexpr int RATE = SNARKV_POSEIDON_T - 1;
#pragma unroll
    for (int q = 0; q < RATE; ++q) {
        if ((uint32_t)q < cnt) {
            Fr e = fp_load<FR>(elems + (size_t)q * 32);
            if (format == SNARKV_CANONICAL) e = fp_to_mont(e);
            s[1 + q] = fp_add(s[1 + q], e);
        } else if ((uint32_t)q == cnt) {
            s[1 + q] = fp_add(s[1 + q], fp_one<FR>());
        }
    }
    poseidon_permute(s);
}

__global__ void __launch_bounds__(64) k_poseidon_transcript(const uint8_t* __restrict__ elements, size_t stream_len, const uint32_t* __restrict__ seg_end,
                                                            uint32_t k, size_t m, int format, uint8_t* __restrict__ out) {
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    constexpr uint32_t RATE = SNARKV_POSEIDON_T - 1;
    const uint8_t* st = elements + j * stream_len * 32;
    Fr s[SNARKV_POSEIDON_T];
    {
        constexpr uint32_t cap[8] = SNARKV_POSEIDON_CAPACITY;
#pragma unroll
        for (int i = 0; i < 8; ++i) s[0].v[i] = cap[i];
#pragma unroll
        for (int i = 1; i < SNARKV_POSEIDON_T; ++i) s[i] = fp_zero<FR>();
    }
    uint32_t prev = 0;
    for (uint32_t i = 0; i < k; ++i) {
        const uint32_t e = seg_end[i];
        for (uint32_t pos = prev; pos < e; pos += RATE) poseidon_absorb(s, st + (size_t)pos * 32, min(RATE, e - pos), format);
        if ((e - prev) % RATE == 0) poseidon_absorb(s, st, 0, format);      // `exact`: one more permutation on an empty chunk
        prev = e;
        Fr c = s[1];
        if (format == SNARKV_CANONICAL) c = fp_from_mont(c);
        fp_store<FR>(out + (j * k + i) * 32, c);
    }
}

// The same transcript with the state spread over lanes: one proof per group of FIVE lanes (six proofs per warp, lanes 30 / 31 shadow
// lanes 25 / 26), lane g holds state[g].  One thread per proof leaves the SMs almost idle at a few thousand proofs (11 ms for 4096 proofs
// of ~60 elements); the first lane version (groups of 8, S-box on every lane, one MDS row per lane: 8 multiplication steps per round)
// took 4.1 ms and was bound by the multiplier SLOTS it occupies (8 lanes x 8 steps x 68 rounds per permutation).  Here a full round is
// still 3 (S-box) + 5 (MDS row) steps, but a PARTIAL round — 60 of the 68 — takes 6: while lane 0 runs its S-box (x^2, x^4, x^5), lanes
// 1..4 multiply what does not depend on it — lane j first the entry M[0][j] x_j of lane 0's row, then M[j][1..4] x_1..4 of its own —
// and one last step multiplies column 0 by the fresh x_0^5 on every lane: 5 lanes x (60 x 6 + 8 x 8) slots per permutation, 2.05 x fewer.
__device__ __forceinline__ Fr fr_shfl(const Fr& a, int src) {
    Fr r;
#pragma unroll
    for (int k = 0; k < 8; ++k) r.v[k] = __shfl_sync(0xffffffffu, a.v[k], src);
    return r;
}
__device__ __forceinline__ Fr fr_sel(bool c, const Fr& a, const Fr& b) {
    Fr r;
#pragma unroll
    for (int k = 0; k < 8; ++k) r.v[k] = c ? a.v[k] : b.v[k];
    return r;
}
__device__ void poseidon_permute_lanes(Fr& s, uint32_t g, int base) {
    constexpr int T = SNARKV_POSEIDON_T, RF = SNARKV_POSEIDON_RF, RP = SNARKV_POSEIDON_RP;
    const bool z = g == 0;
#pragma unroll 1
    for (int r = 0; r < RF + RP; ++r) {
        const Fr x = fp_add(s, fr_const(POSEIDON_RC[r * T + g]));
        if (r < RF / 2 || r >= RF / 2 + RP) {   // full round (warp-uniform)
            const Fr y = fr_pow5(x);
            Fr acc = fp_mul(fr_const(POSEIDON_MDS[g * T]), fr_shfl(y, base));
#pragma unroll
            for (int j = 1; j < T; ++j) acc = fp_add(acc, fp_mul(fr_const(POSEIDON_MDS[g * T + j]), fr_shfl(y, base + j)));
            s = acc;
        } else {                                // partial round: only x_0 goes through the S-box
            const Fr x1 = fr_shfl(x, base + 1), x2 = fr_shfl(x, base + 2), x3 = fr_shfl(x, base + 3), x4 = fr_shfl(x, base + 4);
            // step 1   lane 0: x^2            lane j: M[0][j] x_j  (an entry of lane 0's row)
            const Fr t1 = fp_mul(fr_sel(z, x, fr_const(POSEIDON_MDS[g])), x);
            // step 2   lane 0: x^4            lane j: M[j][1] x_1
            const Fr t2 = fp_mul(fr_sel(z, t1, fr_const(POSEIDON_MDS[g * T + 1])), fr_sel(z, t1, x1));
            // step 3   lane 0: x^5            lane j: M[j][2] x_2
            const Fr t3 = fp_mul(fr_sel(z, t2, fr_const(POSEIDON_MDS[g * T + 2])), fr_sel(z, x, x2));
            // steps 4, 5   lane j: M[j][3] x_3, M[j][4] x_4  (lane 0 idles along)
            const Fr t4 = fp_mul(fr_const(POSEIDON_MDS[g * T + 3]), x3);
            const Fr t5 = fp_mul(fr_const(POSEIDON_MDS[g * T + 4]), x4);
            // step 6   every lane: M[g][0] x_0^5
            const Fr q = fp_mul(fr_const(POSEIDON_MDS[g * T]), fr_shfl(t3, base));
            const Fr row0 = fp_add(fp_add(fr_shfl(t1, base + 1), fr_shfl(t1, base + 2)), fp_add(fr_shfl(t1, base + 3), fr_shfl(t1, base + 4)));
            const Fr own = fp_add(fp_add(t2, t3), fp_add(t4, t5));
            s = fp_add(q, fr_sel(z, row0, own));
        }
    }
}
#define SNARKV_POSEIDON_GROUPS_PER_WARP 6
__global__ void __launch_bounds__(128) k_poseidon_transcript_lanes(const uint8_t* __restrict__ elements, size_t stream_len,
                                                                   const uint32_t* __restrict__ seg_end, uint32_t k, size_t m, int format,
                                                                   uint8_t* __restrict__ out) {
    const uint32_t lane = threadIdx.x & 31u;
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const bool shadow = lane >= 30u;                          // lanes 30, 31 replay lanes 25, 26 (shuffles need every lane) and store nothing
    const uint32_t grp = shadow ? 5u : lane / 5u;
    const uint32_t g = shadow ? lane - 30u : lane - grp * 5u;
    const int base = (int)(grp * 5u);
    const size_t jj = warp * SNARKV_POSEIDON_GROUPS_PER_WARP + grp;
    const size_t j = jj < m ? jj : m - 1;                     // surplus groups of the last warp replay the last proof
    constexpr uint32_t RATE = SNARKV_POSEIDON_T - 1;
    const uint8_t* st = elements + j * stream_len * 32;
    Fr s = fp_zero<FR>();
    if (g == 0) {
        constexpr uint32_t cap[8] = SNARKV_POSEIDON_CAPACITY;
#pragma unroll
        for (int i = 0; i < 8; ++i) s.v[i] = cap[i];
    }
    uint32_t prev = 0;
    for (uint32_t i = 0; i < k; ++i) {
        const uint32_t e = seg_end[i];
        // chunks of RATE elements, then one empty chunk when the count is a multiple of RATE (incl. zero)
        const uint32_t len = e - prev, chunks = len / RATE + 1;
        for (uint32_t c = 0; c < chunks; ++c) {
            const uint32_t pos = prev + c * RATE, cnt = min(RATE, e - pos);
            if (c == chunks - 1 && len % RATE != 0 && cnt == 0) break;
            if (c == chunks - 1 && len % RATE == 0) {
                if (g == 1) s = fp_add(s, fp_one<FR>());                            // the `exact` permutation on an empty chunk
            } else if (g >= 1) {
                if (g - 1 < cnt) {
                    Fr v = fp_load<FR>(st + (size_t)(pos + g - 1) * 32);
                    if (format == SNARKV_CANONICAL) v = fp_to_mont(v);
                    s = fp_add(s, v);
                } else if (g - 1 == cnt) {
                    s = fp_add(s, fp_one<FR>());
                }
            }
            poseidon_permute_lanes(s, g, base);
        }
        prev = e;
        if (g == 1 && !shadow && jj < m) {
            Fr c = s;
            if (format == SNARKV_CANONICAL) c = fp_from_mont(c);
            fp_store<FR>(out + (j * k + i) * 32, c);
        }
    }
}

// parity entry: m states of T elements -> permuted states
__global__ void __launch_bounds__(64) k_poseidon_permute(const uint8_t* __restrict__ in, size_t m, int format, uint8_t* __restrict__ out) {
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    Fr s[SNARKV_POSEIDON_T];
#pragma unroll
    for (int i = 0; i < SNARKV_POSEIDON_T; ++i) {
        s[i] = fp_load<FR>(in + (j * SNARKV_POSEIDON_T + i) * 32);
        if (format == SNARKV_CANONICAL) s[i] = fp_to_mont(s[i]);
    }
    poseidon_permute(s);
#pragma unroll
    for (int i = 0; i < SNARKV_POSEIDON_T; ++i) {
        Fr c = s[i];
        if (format == SNARKV_CANONICAL) c = fp_from_mont(c);
        fp_store<FR>(out + (j * SNARKV_POSEIDON_T + i) * 32, c);
    }
}

// `C::from_bytes` for bn256 G1 (halo2curves): 32 bytes = x little-endian, bit 255 = parity of y, bit 254 = identity flag.
// valid[i] = 1 iff the encoding is a canonical x with a point on the curve; the identity is reported INVALID, as the reference's
// transcript rejects it one line later (`coordinates()` of the identity is None, halo2.rs:226-241) — so the two historical
// layouts of the identity (all zeros / flag bit) need not be told apart.  points: affine (x, y) in `format`;
// elements (optional): x mod r, y mod r in `format` — what common_ec_point absorbs (fe_to_fe).
// one compressed point -> (ok, canonical x / y, Montgomery x / y)
__device__ bool g1_decompress_one(const uint8_t* compressed, Fq& xc, Fq& yc, Fq& xm, Fq& y) {
    Fq x = fp_load<FQ>(compressed);
    const uint32_t sign = x.v[7] >> 31, ident = (x.v[7] >> 30) & 1u;
    x.v[7] &= 0x3fffffffu;
    bool ok = !ident && fp_is_canonical(x);
    xc = x;
    xm = fp_to_mont(x);
    Fq b = fp_one<FQ>();
#pragma unroll
    for (int k = 1; k < SNARKV_CURVE_B; ++k) b = fp_add(b, fp_one<FQ>());
    const Fq rhs = fp_add(fp_mul(fp_sqr(xm), xm), b);
    // y = rhs^((p + 1) / 4)  (p = 3 mod 4)
    uint32_t e[8];
    uint32_t carry = 1;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const uint64_t t = (uint64_t)fp_mod_limb<FQ>(k) + carry;
        e[k] = (uint32_t)t;
        carry = (uint32_t)(t >> 32);
    }
#pragma unroll
    for (int k = 0; k < 7; ++k) e[k] = (e[k] >> 2) | (e[k + 1] << 30);
    e[7] >>= 2;
    y = fp_one<FQ>();
#pragma unroll 1
    for (int bit = 253; bit >= 0; --bit) {
        y = fp_sqr(y);
        if ((e[bit >> 5] >> (bit & 31)) & 1u) y = fp_mul(y, rhs);
    }
    ok = ok && fp_eq(fp_sqr(y), rhs);
    yc = fp_from_mont(y);
    if ((yc.v[0] & 1u) != sign) {
        y = fp_neg(y);
        yc = fp_from_mont(y);
    }
    return ok && !(fp_is_zero(yc) && sign);                      // y = 0 has no odd twin
}
// fe_to_fe: a canonical base-field value reduced into the scalar field (p < 2 r: one conditional subtraction), in `format`
__device__ __forceinline__ Fr fq_to_fr_element(const Fq& c, int format) {
    Fr v;
#pragma unroll
    for (int k = 0; k < 8; ++k) v.v[k] = c.v[k];
    if (!fp_is_canonical(v)) {
        uint32_t borrow = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint64_t d = (uint64_t)v.v[k] - fp_mod_limb<FR>(k) - borrow;
            v.v[k] = (uint32_t)d;
            borrow = (uint32_t)(d >> 63);
        }
    }
    if (format == SNARKV_MONTGOMERY) v = fp_to_mont(v);
    return v;
}

__global__ void __launch_bounds__(128) k_g1_decompress(const uint8_t* __restrict__ compressed, size_t n, int format, uint8_t* __restrict__ points,
                                                       uint8_t* __restrict__ elements, uint8_t* __restrict__ valid) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fq xc, yc, xm, y;
    const bool ok = g1_decompress_one(compressed + i * 32, xc, yc, xm, y);
    if (!ok) {
        valid[i] = 0;
        fp_store<FQ>(points + i * 64, fp_zero<FQ>());
        fp_store<FQ>(points + i * 64 + 32, fp_zero<FQ>());
        if (elements) {
            fp_store<FQ>(elements + i * 64, fp_zero<FQ>());
            fp_store<FQ>(elements + i * 64 + 32, fp_zero<FQ>());
        }
        return;
    }
    valid[i] = 1;
    fp_store<FQ>(points + i * 64, format == SNARKV_CANONICAL ? xc : xm);
    fp_store<FQ>(points + i * 64 + 32, format == SNARKV_CANONICAL ? yc : y);
    if (elements) {
        fp_store<FR>(elements + i * 64, fq_to_fr_element(xc, format));
        fp_store<FR>(elements + i * 64 + 32, fq_to_fr_element(yc, format));
    }
}

// PlonkProof::read over the Poseidon transcript for a batch (plonk_batch.cu): item i of proof j (32 bytes) is a little-endian scalar
// (copied to element `off`, rejected when >= r) or a compressed point (decompressed: the two elements it absorbs go to `off`, `off + 1`,
// the affine point to slot pt_index of the proof's point array).  item_off[i] < 0 marks a point: off = -(item_off[i] + 1).
__global__ void __launch_bounds__(128) k_plonk_poseidon_expand(const uint8_t* __restrict__ proofs, uint32_t n_items, const int32_t* __restrict__ item_off,
                                                               const int32_t* __restrict__ item_pt, uint32_t stream_len, uint32_t n_points, size_t m,
                                                               uint8_t* __restrict__ elements, uint8_t* __restrict__ points, int* __restrict__ status,
                                                               uint32_t in_stride, uint32_t in_base) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m * n_items) return;
    const size_t j = t / n_items;
    const uint32_t i = (uint32_t)(t - j * n_items);
    const uint8_t* src = proofs + (j * in_stride + in_base + i) * 32;   // proof j's items follow its in_base caller-supplied elements
    const int32_t o = item_off[i];
    if (o >= 0) {
        const Fr v = fp_load<FR>(src);
        if (!fp_is_canonical(v)) atomicCAS(status, 0, SNARKV_ERR_BAD_SCALAR);       // "Invalid scalar encoding in proof"
        fp_store<FR>(elements + (j * stream_len + (uint32_t)o) * 32, v);
        return;
    }
    const uint32_t off = (uint32_t)(-(o + 1));
    Fq xc, yc, xm, y;
    if (!g1_decompress_one(src, xc, yc, xm, y)) {
        atomicCAS(status, 0, SNARKV_ERR_BAD_POINT);                                // "Invalid elliptic curve point encoding in proof"
        xc = fp_zero<FQ>();
        yc = fp_zero<FQ>();
    }
    fp_store<FR>(elements + (j * stream_len + off) * 32, fq_to_fr_element(xc, SNARKV_CANONICAL));
    fp_store<FR>(elements + (j * stream_len + off + 1) * 32, fq_to_fr_element(yc, SNARKV_CANONICAL));
    uint8_t* dst = points + (j * n_points + (uint32_t)item_pt[i]) * 64;
    fp_store<FQ>(dst, xc);
    fp_store<FQ>(dst + 32, yc);
}

int plonk_poseidon_expand_device(snarkv_ctx* ctx, const void* d_proofs, uint32_t n_items, const void* d_item_off, const void* d_item_pt, uint32_t stream_len,
                                 uint32_t n_points, size_t m, void* d_elements, void* d_points, void* d_status, uint32_t in_stride, uint32_t in_base) {
    Stage sg(ctx, "plonk_poseidon_expand");
    const size_t n = m * n_items;
    k_plonk_poseidon_expand<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>((const uint8_t*)d_proofs, n_items, (const int32_t*)d_item_off,
                                                                                 (const int32_t*)d_item_pt, stream_len, n_points, m, (uint8_t*)d_elements,
                                                                                 (uint8_t*)d_points, (int*)d_status, in_stride, in_base);
    SNARKV_LAUNCH_CHECK(ctx, "k_plonk_poseidon_expand");
    sg.launched();
    return SNARKV_OK;
}

int poseidon_transcript_device(snarkv_ctx* ctx, const void* d_elements, size_t stream_len, const void* d_seg_end, size_t k, size_t m, int format, void* d_out) {
    Stage sg(ctx, "poseidon_transcript");
    // small and medium batches: one proof per 5 lanes (latency); huge batches: one proof per thread (throughput)
    if (m <= ((size_t)1 << 16))
        k_poseidon_transcript_lanes<<<(unsigned)((m + 4 * SNARKV_POSEIDON_GROUPS_PER_WARP - 1) / (4 * SNARKV_POSEIDON_GROUPS_PER_WARP)), 128, 0, ctx->stream>>>((const uint8_t*)d_elements, stream_len, (const uint32_t*)d_seg_end,
                                                                                            (uint32_t)k, m, format, (uint8_t*)d_out);
    else
        k_poseidon_transcript<<<(unsigned)((m + 63) / 64), 64, 0, ctx->stream>>>((const uint8_t*)d_elements, stream_len, (const uint32_t*)d_seg_end, (uint32_t)k, m,
                                                                               format, (uint8_t*)d_out);
    SNARKV_LAUNCH_CHECK(ctx, "k_poseidon_transcript");
    sg.launched();
    return SNARKV_OK;
}

}  // namespace snarkv

using namespace snarkv;

#define QCTX_GUARD(ctx)                                                                  \
    do {                                                                                 \
        if (!(ctx)) return SNARKV_ERR_USAGE;                                             \
        (ctx)->err.clear();                                                              \
        cudaError_t _g = cudaSetDevice((ctx)->device);                                   \
        if (_g != cudaSuccess) return (ctx)->fail(SNARKV_ERR_CUDA, "cudaSetDevice", _g); \
    } while (0)
static int pos_bad_format(int f) { return f != SNARKV_CANONICAL && f != SNARKV_MONTGOMERY; }

extern "C" {

int snarkv_poseidon_transcript_challenges(snarkv_ctx* ctx, const uint8_t* elements, size_t stream_len, const uint32_t* seg_end, size_t k, size_t m,
                                          int format, uint8_t* challenges) {
    QCTX_GUARD(ctx);
    if (m == 0 || k == 0) return SNARKV_OK;
    if (!seg_end || !challenges || (!elements && stream_len) || pos_bad_format(format))
        return ctx->fail(SNARKV_ERR_USAGE, "snarkv_poseidon_transcript_challenges: bad argument");
    uint32_t prev = 0;
    for (size_t i = 0; i < k; ++i) {
        if (seg_end[i] < prev || seg_end[i] > stream_len) return ctx->fail(SNARKV_ERR_USAGE, "seg_end must be non-decreasing element offsets within the stream");
        prev = seg_end[i];
    }
    ctx->profile_begin_call();
    uint8_t* d_st = (uint8_t*)ctx->wsget(WS_IO_A, m * stream_len * 32 + 32);
    uint8_t* d_seg = (uint8_t*)ctx->wsget(WS_IO_C, k * 4);
    uint8_t* d_out = (uint8_t*)ctx->wsget(WS_IO_B, m * k * 32);
    if (!d_st || !d_seg || !d_out) return SNARKV_ERR_CUDA;
    cudaStream_t st = ctx->stream;
    if (stream_len) SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(d_st, elements, m * stream_len * 32, cudaMemcpyHostToDevice, st));
    SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(d_seg, seg_end, k * 4, cudaMemcpyHostToDevice, st));
    {
        int rc = poseidon_transcript_device(ctx, d_st, stream_len, d_seg, k, m, format, d_out);
        if (rc) return rc;
    }
    SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(challenges, d_out, m * k * 32, cudaMemcpyDeviceToHost, st));
    SNARKV_CUDA_TRY(ctx, cudaStreamSynchronize(st));
    return SNARKV_OK;
}

int snarkv_poseidon_permute(snarkv_ctx* ctx, const uint8_t* states, size_t m, int format, uint8_t* out) {
    QCTX_GUARD(ctx);
    if (m == 0) return SNARKV_OK;
    if (!states || !out || pos_bad_format(format)) return ctx->fail(SNARKV_ERR_USAGE, "snarkv_poseidon_permute: bad argument");
    const size_t bytes = m * SNARKV_POSEIDON_T * 32;
    uint8_t* d_in = (uint8_t*)ctx->wsget(WS_IO_A, bytes);
    uint8_t* d_out = (uint8_t*)ctx->wsget(WS_IO_B, bytes);
    if (!d_in || !d_out) return SNARKV_ERR_CUDA;
    cudaStream_t st = ctx->stream;
    SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(d_in, states, bytes, cudaMemcpyHostToDevice, st));
    k_poseidon_permute<<<(unsigned)((m + 63) / 64), 64, 0, st>>>(d_in, m, format, d_out);
    SNARKV_LAUNCH_CHECK(ctx, "k_poseidon_permute");
    ctx->launches++;
    SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(out, d_out, bytes, cudaMemcpyDeviceToHost, st));
    SNARKV_CUDA_TRY(ctx, cudaStreamSynchronize(st));
    return SNARKV_OK;
}

int snarkv_g1_decompress_batch(snarkv_ctx* ctx, const uint8_t* compressed, size_t n, int format, uint8_t* points, uint8_t* fr_elements, uint8_t* valid) {
    QCTX_GUARD(ctx);
    if (n == 0) return SNARKV_OK;
    if (!compressed || !points || !valid || pos_bad_format(format)) return ctx->fail(SNARKV_ERR_USAGE, "snarkv_g1_decompress_batch: bad argument");
    ctx->profile_begin_call();
    uint8_t* d_in = (uint8_t*)ctx->wsget(WS_IO_A, n * 32);
    uint8_t* d_pts = (uint8_t*)ctx->wsget(WS_IO_B, n * 128);
    uint8_t* d_valid = (uint8_t*)ctx->wsget(WS_IO_C, n);
    if (!d_in || !d_pts || !d_valid) return SNARKV_ERR_CUDA;
    uint8_t* d_el = fr_elements ? d_pts + n * 64 : nullptr;
    cudaStream_t st = ctx->stream;
    SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(d_in, compressed, n * 32, cudaMemcpyHostToDevice, st));
    {
        Stage sg(ctx, "g1_decompress");
        k_g1_decompress<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(d_in, n, format, d_pts, d_el, d_valid);
        SNARKV_LAUNCH_CHECK(ctx, "k_g1_decompress");
        sg.launched();
    }
    SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(points, d_pts, n * 64, cudaMemcpyDeviceToHost, st));
    if (fr_elements) SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(fr_elements, d_el, n * 64, cudaMemcpyDeviceToHost, st));
    SNARKV_CUDA_TRY(ctx, cudaMemcpyAsync(valid, d_valid, n, cudaMemcpyDeviceToHost, st));
    SNARKV_CUDA_TRY(ctx, cudaStreamSynchronize(st));
    return SNARKV_OK;
}

}  // extern "C"
