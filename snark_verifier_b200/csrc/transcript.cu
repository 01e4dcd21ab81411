// transcript.cu — batched Keccak-256 `EvmTranscript` challenges (the second "next" row).
//
// Replaces, for a batch of proofs that share one transcript shape, the native-loader transcript of
//   snark-verifier/src/system/halo2/transcript/evm.rs:175-222   (common_scalar / common_ec_point / squeeze_challenge)
//   snark-verifier/src/loader/evm/util.rs:61-67                  (u256_to_fe: big-endian hash mod r)
// Everything the verifier absorbs is a 32-byte big-endian word, so a proof's transcript is its absorbed byte stream cut at the
// squeeze points:   H_i = Keccak256( H_{i-1} || stream[seg_{i-1} .. seg_i)  [|| 0x01 if that is exactly 32 bytes] ),
// challenge_i = be(H_i) mod r.   Fiat-Shamir is sequential inside one proof and independent across proofs: one thread per proof.
// Keccak-256 is the original padding (0x01 ... 0x80), rate 136 bytes = 17 lanes; all state indices are compile-time constants.
#include "ctx.hpp"
#include "fp.cuh"

namespace snarkv {

__device__ __forceinline__ uint64_t rol64(uint64_t v, int n) { return n ? (v << n) | (v >> (64 - n)) : v; }

__device__ void keccak_f1600(uint64_t s[25]) {
    const uint64_t RC[24] = {0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808Aull, 0x8000000080008000ull,
                             0x000000000000808Bull, 0x0000000080000001ull, 0x8000000080008081ull, 0x8000000000008009ull,
                             0x000000000000008Aull, 0x0000000000000088ull, 0x0000000080008009ull, 0x000000008000000Aull,
                             0x000000008000808Bull, 0x800000000000008Bull, 0x8000000000008089ull, 0x8000000000008003ull,
                             0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800Aull, 0x800000008000000Aull,
                             0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};
    constexpr int ROT[5][5] = {{0, 36, 3, 41, 18}, {1, 44, 10, 45, 2}, {62, 6, 43, 15, 61}, {28, 55, 25, 21, 56}, {27, 20, 39, 8, 14}};
#pragma unroll 1
    for (int rnd = 0; rnd < 24; ++rnd) {
        uint64_t c[5], d[5], b[25];
#pragma unroll
        for (int x = 0; x < 5; ++x) c[x] = s[x] ^ s[x + 5] ^ s[x + 10] ^ s[x + 15] ^ s[x + 20];
#pragma unroll
        for (int x = 0; x < 5; ++x) d[x] = c[(x + 4) % 5] ^ rol64(c[(x + 1) % 5], 1);
#pragma unroll
        for (int x = 0; x < 5; ++x)
#pragma unroll
            for (int y = 0; y < 5; ++y) b[y + 5 * ((2 * x + 3 * y) % 5)] = rol64(s[x + 5 * y] ^ d[x], ROT[x][y]);
#pragma unroll
        for (int y = 0; y < 5; ++y)
#pragma unroll
            for (int x = 0; x < 5; ++x) s[x + 5 * y] = b[x + 5 * y] ^ (~b[(x + 1) % 5 + 5 * y] & b[(x + 2) % 5 + 5 * y]);
        s[0] ^= RC[rnd];
    }
}

__global__ void __launch_bounds__(128) k_evm_transcript(const uint8_t* __restrict__ streams, size_t stream_len,
                                                        const uint32_t* __restrict__ seg_end, uint32_t k, size_t m, int format,
                                                        uint8_t* __restrict__ out) {
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const uint64_t* st = reinterpret_cast<const uint64_t*>(streams + j * stream_len);   // 8-byte lanes, little-endian loads
    uint64_t h0 = 0, h1 = 0, h2 = 0, h3 = 0;
    uint32_t prev = 0;
    for (uint32_t i = 0; i < k; ++i) {
        const uint32_t e = seg_end[i];
        const uint32_t hl = i ? 4u : 0u;                        // lanes of the previous hash that start the buffer
        const uint32_t lanes = hl + (e - prev) / 8u;            // message length in whole lanes
        const uint64_t pad_lane = (lanes == 4u) ? 0x0101ull : 0x01ull;   // buf.len() == 0x20 -> data byte 0x01, then the pad byte
        const uint32_t blocks = lanes / 17u + 1u;
        uint64_t s[25];
#pragma unroll
        for (int t = 0; t < 25; ++t) s[t] = 0;
        for (uint32_t b = 0; b < blocks; ++b) {
#pragma unroll
            for (int t = 0; t < 17; ++t) {
                const uint32_t q = b * 17u + (uint32_t)t;
                uint64_t lane = 0;
                if (q < lanes) {
                    if (q < hl) lane = q == 0 ? h0 : q == 1 ? h1 : q == 2 ? h2 : h3;
                    else lane = st[prev / 8u + (q - hl)];
                } else if (q == lanes) lane = pad_lane;
                if (t == 16 && b == blocks - 1u) lane ^= 0x8000000000000000ull;
                s[t] ^= lane;
            }
            keccak_f1600(s);
        }
        h0 = s[0]; h1 = s[1]; h2 = s[2]; h3 = s[3];
        prev = e;
        // challenge = U256::from_be_bytes(hash) % r : the hash bytes are the little-endian bytes of h0..h3 in order
        uint64_t v[4] = {__byte_perm((uint32_t)(h3 >> 32), 0, 0x0123) | ((uint64_t)__byte_perm((uint32_t)h3, 0, 0x0123) << 32),
                         __byte_perm((uint32_t)(h2 >> 32), 0, 0x0123) | ((uint64_t)__byte_perm((uint32_t)h2, 0, 0x0123) << 32),
                         __byte_perm((uint32_t)(h1 >> 32), 0, 0x0123) | ((uint64_t)__byte_perm((uint32_t)h1, 0, 0x0123) << 32),
                         __byte_perm((uint32_t)(h0 >> 32), 0, 0x0123) | ((uint64_t)__byte_perm((uint32_t)h0, 0, 0x0123) << 32)};
        Fr c;
#pragma unroll
        for (int t = 0; t < 4; ++t) { c.v[2 * t] = (uint32_t)v[t]; c.v[2 * t + 1] = (uint32_t)(v[t] >> 32); }
        for (int it = 0; it < 6 && !fp_is_canonical(c); ++it) {   // 2^256 / r < 6
            uint32_t borrow = 0;
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const uint64_t dd = (uint64_t)c.v[t] - fp_mod_limb<FR>(t) - borrow;
                c.v[t] = (uint32_t)dd;
                borrow = (uint32_t)(dd >> 63);
            }
        }
        if (format == SNARKV_MONTGOMERY) c = fp_to_mont(c);
        fp_store<FR>(out + (j * k + i) * 32, c);
    }
}

int evm_transcript_device(snarkv_ctx* ctx, const void* d_streams, size_t stream_len, const void* d_seg_end, size_t k, size_t m, int format,
                          void* d_out) {
    Stage sg(ctx, "evm_transcript");
    k_evm_transcript<<<(unsigned)((m + 127) / 128), 128, 0, ctx->stream>>>((const uint8_t*)d_streams, stream_len, (const uint32_t*)d_seg_end,
                                                                        (uint32_t)k, m, format, (uint8_t*)d_out);
    SNARKV_LAUNCH_CHECK(ctx, "k_evm_transcript");
    sg.launched();
    return SNARKV_OK;
}

}  // namespace snarkv
