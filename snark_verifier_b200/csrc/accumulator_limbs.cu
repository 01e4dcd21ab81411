// accumulator_limbs.cu — `LimbsEncoding<LIMBS, BITS>::from_repr` for a batch of accumulators.
//
// Replaces snark-verifier/src/pcs/kzg/accumulator.rs:57-81 (native `AccumulatorEncoding::from_repr`) with
// snark-verifier/src/util/arithmetic.rs:270-282 (`fe_from_limbs`): an old KZG accumulator travels through a proof's instances as
// 4 x LIMBS scalar-field limbs (lhs.x, lhs.y, rhs.x, rhs.y; BITS bits each, 4 x 68 in the SDK), and the verifier rebuilds the two
// G1 points before it hands them to `decide_all`:
//     coordinate = sum_i limb_i << (BITS i)      (as integers; must fit 32 bytes and be a canonical base-field element)
//     point      = C::from_xy(x, y).unwrap()     (on the curve, or the identity encoded as (0, 0))
// The reference panics on a violation; here it is data: valid[a] = 0.  One thread per accumulator.
#include "ctx.hpp"
#include "g1.cuh"

namespace snarkv {

// acc (16 words, little endian) += v << shift; returns false if anything lands at or above bit 256
__device__ __forceinline__ bool add_shifted(uint32_t acc[16], const Fr& v, uint32_t shift) {
    const uint32_t ws = shift >> 5, bs = shift & 31u;
    uint64_t carry = 0;
#pragma unroll 1
    for (uint32_t i = 0; i < 9; ++i) {
        const uint32_t lo = i < 8 ? v.v[i] : 0u, prev = i > 0 ? v.v[i - 1] : 0u;
        const uint32_t part = bs ? ((lo << bs) | (prev >> (32u - bs))) : lo;
        const uint32_t k = ws + i;
        if (k < 16) {
            const uint64_t t = (uint64_t)acc[k] + part + carry;
            acc[k] = (uint32_t)t;
            carry = t >> 32;
        } else if (part || carry) return false;
    }
    for (uint32_t k = ws + 9; carry && k < 16; ++k) {
        const uint64_t t = (uint64_t)acc[k] + carry;
        acc[k] = (uint32_t)t;
        carry = t >> 32;
    }
    return carry == 0;
}

__global__ void __launch_bounds__(128) k_accumulators_from_limbs(const uint8_t* __restrict__ limbs, size_t m, uint32_t L, uint32_t bits,
                                                                 int format, uint8_t* __restrict__ lhs, uint8_t* __restrict__ rhs,
                                                                 uint8_t* __restrict__ valid) {
    const size_t a = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= m) return;
    bool ok = true;
    Fq coord[4];
    for (uint32_t c = 0; c < 4; ++c) {
        uint32_t acc[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) acc[k] = 0;
        for (uint32_t i = 0; i < L; ++i) {
            Fr v = fp_load<FR>(limbs + ((a * 4 + c) * L + i) * 32);
            if (format == SNARKV_MONTGOMERY) v = fp_from_mont(v);
            else ok = ok && fp_is_canonical(v);   // a limb is an Fr value: `from_repr` of a non-canonical one fails upstream
            ok = add_shifted(acc, v, bits * i) && ok;
        }
        uint32_t high = 0;
#pragma unroll
        for (int k = 8; k < 16; ++k) high |= acc[k];
        Fq x;
#pragma unroll
        for (int k = 0; k < 8; ++k) x.v[k] = acc[k];
        ok = ok && high == 0 && fp_is_canonical(x);
        coord[c] = fp_to_mont(x);
    }
    G1Affine p0{coord[0], coord[1]}, p1{coord[2], coord[3]};
    ok = ok && g1_affine_is_on_curve(p0) && g1_affine_is_on_curve(p1);
    if (!ok) {   // never hand a non-point downstream
        p0.x = p0.y = p1.x = p1.y = fp_zero<FQ>();
    } else if (format == SNARKV_CANONICAL) {
        p0.x = fp_from_mont(p0.x); p0.y = fp_from_mont(p0.y);
        p1.x = fp_from_mont(p1.x); p1.y = fp_from_mont(p1.y);
    }
    g1_affine_store(lhs, a, p0);
    g1_affine_store(rhs, a, p1);
    valid[a] = ok ? 1 : 0;
}

int accumulators_from_limbs_device(snarkv_ctx* ctx, const void* d_limbs, size_t m, uint32_t L, uint32_t bits, int format, void* d_lhs,
                                   void* d_rhs, void* d_valid) {
    Stage sg(ctx, "accumulators_from_limbs");
    k_accumulators_from_limbs<<<(unsigned)((m + 127) / 128), 128, 0, ctx->stream>>>((const uint8_t*)d_limbs, m, L, bits, format, (uint8_t*)d_lhs,
                                                                                  (uint8_t*)d_rhs, (uint8_t*)d_valid);
    SNARKV_LAUNCH_CHECK(ctx, "k_accumulators_from_limbs");
    sg.launched();
    return SNARKV_OK;
}

}  // namespace snarkv
