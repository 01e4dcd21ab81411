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

}  // namespace snarkv
