// fp_inv.cuh — modular inverse by the binary extended Euclidean algorithm (no multiplications, ~2 x 254 shift/subtract
// steps).  Used where ONE inversion sits on a serial critical path (`Curve::to_affine`, loader/native.rs:70): as a dependent
// chain it is ~5x shorter than the 254-squaring Fermat ladder.  Plain C on 8 x 32-bit limbs, __host__ __device__ so that the
// exact code is unit-tested on the CPU (tests/test_fp_inv_host.py compiles it with g++).
#pragma once
#include <cstdint>

#ifndef __CUDACC__
#define SNARKV_HD
#else
#define SNARKV_HD __host__ __device__ __forceinline__
#endif

namespace snarkv {

struct U256 { uint32_t v[8]; };

SNARKV_HD bool u256_is_even(const U256& a) { return (a.v[0] & 1u) == 0; }
SNARKV_HD bool u256_is_one(const U256& a) {
    uint32_t r = a.v[0] ^ 1u;
    for (int i = 1; i < 8; ++i) r |= a.v[i];
    return r == 0;
}
SNARKV_HD bool u256_is_zero(const U256& a) {
    uint32_t r = 0;
    for (int i = 0; i < 8; ++i) r |= a.v[i];
    return r == 0;
}
SNARKV_HD bool u256_geq(const U256& a, const U256& b) {
    for (int i = 7; i >= 0; --i) {
        if (a.v[i] > b.v[i]) return true;
        if (a.v[i] < b.v[i]) return false;
    }
    return true;
}
SNARKV_HD void u256_shr1(U256& a) {
    for (int i = 0; i < 7; ++i) a.v[i] = (a.v[i] >> 1) | (a.v[i + 1] << 31);
    a.v[7] >>= 1;
}
SNARKV_HD uint32_t u256_add(U256& a, const U256& b) {  // a += b, returns carry
    uint64_t c = 0;
    for (int i = 0; i < 8; ++i) {
        c += (uint64_t)a.v[i] + b.v[i];
        a.v[i] = (uint32_t)c;
        c >>= 32;
    }
    return (uint32_t)c;
}
SNARKV_HD uint32_t u256_sub(U256& a, const U256& b) {  // a -= b, returns borrow
    uint64_t br = 0;
    for (int i = 0; i < 8; ++i) {
        uint64_t d = (uint64_t)a.v[i] - b.v[i] - br;
        a.v[i] = (uint32_t)d;
        br = (d >> 63) & 1u;
    }
    return (uint32_t)br;
}
// x / 2 mod p for odd p < 2^255
SNARKV_HD void u256_half_mod(U256& x, const U256& p) {
    if (!u256_is_even(x)) u256_add(x, p);  // x + p < 2^255: no carry out
    u256_shr1(x);
}
// x - y mod p, inputs < p
SNARKV_HD void u256_sub_mod(U256& x, const U256& y, const U256& p) {
    if (u256_sub(x, y)) u256_add(x, p);
}

// a^-1 mod p for 0 < a < p, p an odd prime < 2^255.  Returns 0 for a == 0.
SNARKV_HD U256 u256_inv_mod(const U256& a, const U256& p) {
    U256 u = a, v = p, b, c;
    for (int i = 0; i < 8; ++i) { b.v[i] = 0; c.v[i] = 0; }
    b.v[0] = 1;
    if (u256_is_zero(u)) return c;
    // invariants: b * a == u (mod p), c * a == v (mod p)
    while (!u256_is_one(u) && !u256_is_one(v)) {
        while (u256_is_even(u)) { u256_shr1(u); u256_half_mod(b, p); }
        while (u256_is_even(v)) { u256_shr1(v); u256_half_mod(c, p); }
        if (u256_geq(u, v)) { u256_sub(u, v); u256_sub_mod(b, c, p); }
        else { u256_sub(v, u); u256_sub_mod(c, b, p); }
    }
    return u256_is_one(u) ? b : c;
}

// -----------------------------------------------------------------------------------------------------------------------
// Faster inverse for the same serial call sites: the binary GCD run 31 steps at a time on 64-bit approximations of (a, b)
// (T. Pornin, "Optimized Binary GCD for Modular Inversion", 2020): an inner loop on two 64-bit words collects the update
// factors f0, g0, f1, g1 (|f| + |g| <= 2^31), which are then applied to the full-width values,
//     (a, b) <- (a f0 + b g0, a f1 + b g1) / 2^31        (exact; a negative result is negated together with its factors)
//     (u, v) <- (u f0 + v g0, u f1 + v g1) / 2^31 mod p  (one 31-bit Montgomery step),
// keeping a = u y, b = v y (mod p).  2 * 254 - 1 = 507 steps suffice for a 254-bit modulus; 17 rounds make 527.  About 4x
// fewer instructions than the bit-at-a-time loop above and no data-dependent trip counts.
// -----------------------------------------------------------------------------------------------------------------------
struct U288 { uint32_t v[9]; };

// r = x * k for a 32-bit k (k <= 2^31)
SNARKV_HD U288 u256_mul_small(const U256& x, uint32_t k) {
    U288 r;
    uint64_t c = 0;
    for (int i = 0; i < 8; ++i) {
        c += (uint64_t)x.v[i] * k;
        r.v[i] = (uint32_t)c;
        c >>= 32;
    }
    r.v[8] = (uint32_t)c;
    return r;
}
SNARKV_HD void u288_add(U288& a, const U288& b) {
    uint64_t c = 0;
    for (int i = 0; i < 9; ++i) {
        c += (uint64_t)a.v[i] + b.v[i];
        a.v[i] = (uint32_t)c;
        c >>= 32;
    }
}
SNARKV_HD uint32_t u288_sub(U288& a, const U288& b) {  // a -= b, returns borrow
    uint64_t br = 0;
    for (int i = 0; i < 9; ++i) {
        uint64_t d = (uint64_t)a.v[i] - b.v[i] - br;
        a.v[i] = (uint32_t)d;
        br = (d >> 63) & 1u;
    }
    return (uint32_t)br;
}
SNARKV_HD void u288_neg(U288& a) {
    uint64_t c = 1;
    for (int i = 0; i < 9; ++i) {
        c += (uint64_t)(~a.v[i]);
        a.v[i] = (uint32_t)c;
        c >>= 32;
    }
}
SNARKV_HD U256 u288_shr31(const U288& a) {  // low 256 bits of a >> 31
    U256 r;
    for (int i = 0; i < 8; ++i) r.v[i] = (a.v[i] >> 31) | (a.v[i + 1] << 1);
    return r;
}
// |x f + y g| for signed factors given as magnitude + sign; returns the sign of x f + y g (1 = negative)
SNARKV_HD uint32_t u256_lincomb(const U256& x, uint32_t fm, uint32_t fs, const U256& y, uint32_t gm, uint32_t gs, U288& out) {
    U288 A = u256_mul_small(x, fm);
    const U288 B = u256_mul_small(y, gm);
    if (fs == gs) {
        u288_add(A, B);
        out = A;
        return fs;
    }
    uint32_t s = fs;
    if (u288_sub(A, B)) {  // |B| > |A|: the sign is g's
        u288_neg(A);
        s = gs;
    }
    out = A;
    return s;
}
// bits [pos, pos + 33) of a, for 32 <= pos and pos + 33 <= 256
SNARKV_HD uint64_t u256_bits33(const U256& a, uint32_t pos) {
    const uint32_t wi = pos >> 5, sh = pos & 31u;
    uint32_t w0 = 0, w1 = 0;
    for (uint32_t i = 1; i < 8; ++i) {  // selects instead of dynamic indexing (registers on the device)
        if (i == wi) w0 = a.v[i];
        if (i == wi + 1) w1 = a.v[i];
    }
    return ((((uint64_t)w1 << 32) | w0) >> sh) & 0x1ffffffffull;
}

// a^-1 mod p for 0 < a < p, p an odd prime < 2^255 (of at most 254 bits for the round count below).  Returns 0 for a == 0.
SNARKV_HD U256 u256_inv_mod_fast(const U256& y, const U256& p) {
    U256 a = y, b = p, u, v;
    for (int i = 0; i < 8; ++i) { u.v[i] = 0; v.v[i] = 0; }
    u.v[0] = 1;
    uint32_t pinv = p.v[0];                                      // -p^-1 mod 2^31 by Newton's iteration
    for (int i = 0; i < 4; ++i) pinv *= 2u - p.v[0] * pinv;
    pinv = (0u - pinv) & 0x7fffffffu;
#ifdef __CUDA_ARCH__
#pragma unroll 1   // compact code: this runs on one lane while other warps need the instruction cache
#endif
    for (int round = 0; round < 17; ++round) {
        // 64-bit approximations: low 31 bits + the 33 bits below the common top bit
        uint32_t top = 0;
        for (uint32_t i = 0; i < 8; ++i)
            if ((a.v[i] | b.v[i]) != 0) top = i;
        uint32_t tw = 0;
        for (uint32_t i = 0; i < 8; ++i)
            if (i == top) tw = a.v[i] | b.v[i];
        uint32_t lz = 0;
        while (lz < 32 && ((tw << lz) & 0x80000000u) == 0) ++lz;
        const uint32_t n = 32u * top + 32u - lz;                 // max bit length (0 never reaches here with b odd >= 1)
        uint64_t xa, xb;
        if (n <= 64) {
            xa = ((uint64_t)a.v[1] << 32) | a.v[0];
            xb = ((uint64_t)b.v[1] << 32) | b.v[0];
        } else {
            xa = (a.v[0] & 0x7fffffffu) | (u256_bits33(a, n - 33) << 31);
            xb = (b.v[0] & 0x7fffffffu) | (u256_bits33(b, n - 33) << 31);
        }
        int64_t f0 = 1, g0 = 0, f1 = 0, g1 = 1;
#ifdef __CUDA_ARCH__
#pragma unroll 1
#endif
        for (int i = 0; i < 31; ++i) {
            if (xa & 1u) {
                if (xa < xb) {
                    const uint64_t tx = xa; xa = xb; xb = tx;
                    int64_t tf = f0; f0 = f1; f1 = tf;
                    tf = g0; g0 = g1; g1 = tf;
                }
                xa -= xb; f0 -= f1; g0 -= g1;
            }
            xa >>= 1;
            f1 <<= 1; g1 <<= 1;
        }
        uint32_t f0s = f0 < 0, g0s = g0 < 0, f1s = f1 < 0, g1s = g1 < 0;
        const uint32_t f0m = (uint32_t)(f0 < 0 ? -f0 : f0), g0m = (uint32_t)(g0 < 0 ? -g0 : g0);
        const uint32_t f1m = (uint32_t)(f1 < 0 ? -f1 : f1), g1m = (uint32_t)(g1 < 0 ? -g1 : g1);
        U288 ta, tb;
        if (u256_lincomb(a, f0m, f0s, b, g0m, g0s, ta)) { f0s ^= 1u; g0s ^= 1u; }   // negative: negate value and factors
        if (u256_lincomb(a, f1m, f1s, b, g1m, g1s, tb)) { f1s ^= 1u; g1s ^= 1u; }
        a = u288_shr31(ta);
        b = u288_shr31(tb);
        // (u, v) <- the same combinations / 2^31 mod p
        U256 nu[2];
        for (int k = 0; k < 2; ++k) {
            U288 t;
            const uint32_t neg = k == 0 ? u256_lincomb(u, f0m, f0s, v, g0m, g0s, t) : u256_lincomb(u, f1m, f1s, v, g1m, g1s, t);
            const uint32_t q = (t.v[0] * pinv) & 0x7fffffffu;    // t + q p = 0 (mod 2^31), t + q p < 2^32 p
            u288_add(t, u256_mul_small(p, q));
            U256 r = u288_shr31(t);                              // < 2 p
            if (u256_geq(r, p)) u256_sub(r, p);
            if (neg && !u256_is_zero(r)) { U256 pr = p; u256_sub(pr, r); r = pr; }
            nu[k] = r;
        }
        u = nu[0];
        v = nu[1];
    }
    if (!u256_is_one(b)) { for (int i = 0; i < 8; ++i) v.v[i] = 0; }
    return v;
}

}  // namespace snarkv
