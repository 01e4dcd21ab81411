// glv.cuh — GLV decomposition of a BN254 scalar for the endomorphism phi(x, y) = (beta x, y) = [lambda] (x, y).
//
// k (canonical, < r)  ->  (k1, k2), |k1|, |k2| < 2^128, with  k = k1 + lambda k2 (mod r), so that
//   k P = |k1| (+-P) + |k2| (+-phi(P)) :  two half-length scalars  =>  half the windows of the bucket method
// (half the Horner chain of util/msm.rs:285-287, half the bucket-reduction work of :298-302).  The affine result of the MSM is
// unchanged, so this is invisible at the C-ABI boundary.
//
// Lattice basis (a1, b1), (a2, b2) of { (a, b) : a + b lambda = 0 mod r }, det = +r, b1 < 0, a2 = |b1| (derived with Python big
// integers by extended Euclid on (r, lambda); every identity is re-asserted in tests/test_glv_host.py):
//   c1 = floor(k g1 / 2^256), g1 = floor(2^256 b2 / r)     c2 = floor(k g2 / 2^256), g2 = floor(2^256 |b1| / r)
//   k1 = k - c1 a1 - c2 a2                                   k2 = c1 |b1| - c2 b2
// Plain C on 32-bit limbs, __host__ __device__: the exact code is unit-tested on the CPU (g++) against the Python formula.
#pragma once
#include <cstdint>

#ifndef __CUDACC__
#define SNARKV_GLV_HD inline
#else
#define SNARKV_GLV_HD __host__ __device__ __forceinline__
#endif

namespace snarkv {
namespace glv {

// lambda = 0x30644e72e131a029048b6e193fd84104cc37a73fec2bc5e9b8ca0b2d36636f23 (cube root of unity in Fr)
// beta   = 0x30644e72e131a0295e6dd9e7e0acccb0c28f069fbb966e3de4bd44e5607cfd48 (cube root of unity in Fq), [lambda](x,y) = (beta x, y)
#define SNARKV_GLV_BETA_MONT_LIMBS {0x13e80b9cu, 0x3350c88eu, 0xdb5e56b9u, 0x7dce557cu, 0xb615564au, 0x6001b4b8u, 0x020217e0u, 0x2682e617u}

// out[0..nout) = low nout limbs of a[0..na) * b[0..nb)
SNARKV_GLV_HD void mul_lo(const uint32_t* a, int na, const uint32_t* b, int nb, uint32_t* out, int nout) {
    for (int i = 0; i < nout; ++i) out[i] = 0;
    for (int i = 0; i < na; ++i) {
        uint64_t carry = 0;
        for (int j = 0; j < nb && i + j < nout; ++j) {
            uint64_t t = (uint64_t)a[i] * b[j] + out[i + j] + carry;
            out[i + j] = (uint32_t)t;
            carry = t >> 32;
        }
        for (int p = i + nb; p < nout && carry; ++p) {
            uint64_t t = (uint64_t)out[p] + carry;
            out[p] = (uint32_t)t;
            carry = t >> 32;
        }
    }
}
SNARKV_GLV_HD void sub6(uint32_t* a, const uint32_t* b) {  // a -= b  (mod 2^192)
    uint64_t br = 0;
    for (int i = 0; i < 6; ++i) {
        uint64_t d = (uint64_t)a[i] - b[i] - br;
        a[i] = (uint32_t)d;
        br = (d >> 63) & 1u;
    }
}
// two's-complement 192-bit value -> magnitude (returns 1 if it was negative)
SNARKV_GLV_HD uint32_t abs6(uint32_t* a) {
    if (!(a[5] >> 31)) return 0;
    uint64_t c = 1;
    for (int i = 0; i < 6; ++i) {
        c += (uint64_t)(~a[i]);
        a[i] = (uint32_t)c;
        c >>= 32;
    }
    return 1;
}

// k: canonical scalar (< r), 8 limbs.  Outputs: magnitudes m1, m2 (5 limbs each, < 2^128 so limb 4 is 0 — kept for the digit
// shifter), signs neg1, neg2 (1 = negative).
SNARKV_GLV_HD void decompose(const uint32_t k[8], uint32_t m1[5], uint32_t& neg1, uint32_t m2[5], uint32_t& neg2) {
    const uint32_t A1[4] = {0x7d4f1128u, 0x8211bbebu, 0xeeb859fcu, 0x6f4d8248u};
    const uint32_t B1ABS[2] = {0x94d213e3u, 0x89d32568u};  // |b1| = a2
    const uint32_t B2[4] = {0x1221250bu, 0x0be4e154u, 0xeeb859fdu, 0x6f4d8248u};
    const uint32_t G1[5] = {0x00ff6565u, 0x5398fd03u, 0xa773d2d2u, 0x4ccef014u, 0x00000002u};
    const uint32_t G2[3] = {0xc7e0b3d7u, 0xd91d232eu, 0x00000002u};
    uint32_t t[13];
    mul_lo(k, 8, G1, 5, t, 13);
    uint32_t c1[5] = {t[8], t[9], t[10], t[11], t[12]};
    mul_lo(k, 8, G2, 3, t, 11);
    uint32_t c2[3] = {t[8], t[9], t[10]};
    uint32_t k1[6], k2[6], u[6];
    for (int i = 0; i < 6; ++i) k1[i] = k[i];
    mul_lo(c1, 5, A1, 4, u, 6);
    sub6(k1, u);
    mul_lo(c2, 3, B1ABS, 2, u, 6);   // c2 * a2
    sub6(k1, u);
    mul_lo(c1, 5, B1ABS, 2, k2, 6);  // c1 * |b1|
    mul_lo(c2, 3, B2, 4, u, 6);
    sub6(k2, u);
    neg1 = abs6(k1);
    neg2 = abs6(k2);
    for (int i = 0; i < 5; ++i) { m1[i] = k1[i]; m2[i] = k2[i]; }
}

}  // namespace glv
}  // namespace snarkv
