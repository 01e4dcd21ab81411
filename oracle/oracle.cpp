// oracle/oracle.cpp — CPU restatement of snark-verifier's native MSM + KZG-decide hot path (TEST INFRASTRUCTURE ONLY;
// see the header of oracle/bn254.hpp for the rules and the "parity unpinned" statement).
//
// Reference algorithms restated here (paths relative to /root/reference/snark-verifier/src):
//   msm_native_fold        <- loader/native.rs:61-71      NativeLoader::multi_scalar_multiplication
//   msm_pippenger_serial   <- util/msm.rs:259-304         multi_scalar_multiplication_serial (+ enum Bucket :228-257)
//   msm_pippenger          <- util/msm.rs:308-343         rayon chunking + fold
//   kzg_decide             <- pcs/kzg/decider.rs:70-82    KzgAs::decide (native)
//   oracle_kzg_decide_all  <- pcs/kzg/decider.rs:84-93    decide_all = loop, first failure aborts
//   oracle_kzg_accumulate  <- pcs/kzg/accumulation.rs:41-63  KzgAs::verify (powers of r, two MSMs)
// Curve/pairing internals restate halo2curves 0.6.0's published algorithms (Jacobian a=0 formulas from the EFD,
// optimal-ate Miller loop with precomputed line coefficients, Devegili/Scott hard-part chain).
#include "bn254.hpp"

#include <cmath>
#include <mutex>
#include <thread>

namespace oracle {

thread_local u64 g_mulmod_count = 0;

const Modulus FQ_MOD = {
    {0x3c208c16d87cfd47ull, 0x97816a916871ca8dull, 0xb85045b68181585dull, 0x30644e72e131a029ull},
    0x87d20782e4866389ull,
    {0xd35d438dc58f0d9dull, 0x0a78eb28f5c70b3dull, 0x666ea36f7879462cull, 0x0e0a77c19a07df2full},
    {0xf32cfc5b538afa89ull, 0xb5e71911d44501fbull, 0x47ab1eff0a417ff6ull, 0x06d89f71cab8351full}};
const Modulus FR_MOD = {
    {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull},
    0xc2e1f593efffffffull,
    {0xac96341c4ffffffbull, 0x36fc76959f60cd29ull, 0x666ea36f7879462eull, 0x0e0a77c19a07df2full},
    {0x1bb8e645ae216da7ull, 0x53fe3ab1e35c59e3ull, 0x8c49833d53bb8085ull, 0x0216d0b17f4e44a5ull}};

static const u64 BN_U = 0x44e992b44a6909f1ull;

// =====================================================================================================================
// G1
// =====================================================================================================================
bool G1Affine::is_on_curve() const {
    if (is_identity()) return true;
    return y.sqr() == x.sqr() * x + Fq::from_u64(3);
}

G1 G1::dbl() const {
    if (is_identity()) return *this;
    Fq a = x.sqr(), b = y.sqr(), c = b.sqr();
    Fq d = ((x + b).sqr() - a - c).dbl();
    Fq e = a.dbl() + a;
    Fq f = e.sqr();
    G1 r;
    r.x = f - d.dbl();
    r.z = (y * z).dbl();
    r.y = e * (d - r.x) - c.dbl().dbl().dbl();
    return r;
}

G1 G1::add(const G1& o) const {
    if (is_identity()) return o;
    if (o.is_identity()) return *this;
    Fq z1z1 = z.sqr(), z2z2 = o.z.sqr();
    Fq u1 = x * z2z2, u2 = o.x * z1z1;
    Fq s1 = y * o.z * z2z2, s2 = o.y * z * z1z1;
    if (u1 == u2) {
        if (s1 == s2) return dbl();
        return identity();
    }
    Fq h = u2 - u1;
    Fq i = h.dbl().sqr();
    Fq j = h * i;
    Fq rr = (s2 - s1).dbl();
    Fq v = u1 * i;
    G1 r;
    r.x = rr.sqr() - j - v.dbl();
    r.y = rr * (v - r.x) - (s1 * j).dbl();
    r.z = ((z + o.z).sqr() - z1z1 - z2z2) * h;
    return r;
}

G1 G1::add_mixed(const G1Affine& o) const {
    if (o.is_identity()) return *this;
    if (is_identity()) return from_affine(o);
    Fq z1z1 = z.sqr();
    Fq u2 = o.x * z1z1;
    Fq s2 = o.y * z * z1z1;
    if (x == u2) {
        if (y == s2) return dbl();
        return identity();
    }
    Fq h = u2 - x;
    Fq hh = h.sqr();
    Fq i = hh.dbl().dbl();
    Fq j = h * i;
    Fq rr = (s2 - y).dbl();
    Fq v = x * i;
    G1 r;
    r.x = rr.sqr() - j - v.dbl();
    r.y = rr * (v - r.x) - (y * j).dbl();
    r.z = (z + h).sqr() - z1z1 - hh;
    return r;
}

G1Affine G1::to_affine() const {
    if (is_identity()) return G1Affine::identity();
    Fq zi = z.inv();
    Fq zi2 = zi.sqr();
    return {x * zi2, y * zi2 * zi};
}

G1 g1_mul_ct(const G1Affine& p, const Fr& s) {
    u64 k[4];
    s.to_raw(k);
    G1 acc = G1::identity();
    for (int i = 254; i >= 0; --i) {  // 256 bits, MSB skipped (skip(1)); bit 255 and 254 of a <2^254 value are 0
        acc = acc.dbl();
        G1 sum = acc.add_mixed(p);  // always computed (conditional_select in the published code)
        if ((k[i / 64] >> (i % 64)) & 1) acc = sum;
    }
    return acc;
}

G1 g1_mul_vartime(const G1Affine& p, const u64 k[4]) {
    G1 acc = G1::identity();
    int top = 255;
    while (top >= 0 && !((k[top / 64] >> (top % 64)) & 1)) --top;
    for (int i = top; i >= 0; --i) {
        acc = acc.dbl();
        if ((k[i / 64] >> (i % 64)) & 1) acc = acc.add_mixed(p);
    }
    return acc;
}

void g1_batch_to_affine(const G1* in, G1Affine* out, size_t n) {
    // Montgomery's trick: one inversion for the batch
    std::vector<Fq> prefix(n);
    Fq acc = Fq::one();
    for (size_t i = 0; i < n; ++i) {
        prefix[i] = acc;
        if (!in[i].is_identity()) acc = acc * in[i].z;
    }
    Fq inv = acc.inv();
    for (size_t i = n; i-- > 0;) {
        if (in[i].is_identity()) { out[i] = G1Affine::identity(); continue; }
        Fq zi = inv * prefix[i];
        inv = inv * in[i].z;
        Fq zi2 = zi.sqr();
        out[i] = {in[i].x * zi2, in[i].y * zi2 * zi};
    }
}

G1Affine msm_native_fold(const Fr* scalars, const G1Affine* bases, size_t n) {
    // pairs.iter().map(|(scalar, base)| *base * scalar).reduce(|acc, value| acc + value).unwrap().to_affine()
    G1 acc = g1_mul_ct(bases[0], scalars[0]);
    for (size_t i = 1; i < n; ++i) acc = acc.add(g1_mul_ct(bases[i], scalars[i]));
    return acc.to_affine();
}

namespace {
// enum Bucket { None, Affine(C), Projective(C::Curve) }   util/msm.rs:228-257
struct Bucket {
    enum Kind : uint8_t { NONE, AFFINE, PROJECTIVE } kind = NONE;
    G1Affine a;
    G1 p;
    void add_assign(const G1Affine& rhs) {
        switch (kind) {
            case NONE: a = rhs; kind = AFFINE; break;
            case AFFINE: p = G1::from_affine(a).add_mixed(rhs); kind = PROJECTIVE; break;
            case PROJECTIVE: p = p.add_mixed(rhs); break;
        }
    }
    G1 add(const G1& rhs) const {
        switch (kind) {
            case NONE: return rhs;
            case AFFINE: return rhs.add_mixed(a);
            default: return p.add(rhs);
        }
    }
};
}  // namespace

void msm_pippenger_serial(const Fr* scalars, const G1Affine* bases, size_t n, G1* result) {
    std::vector<uint8_t> repr(n * 32 + 8, 0);  // to_repr(): canonical LE bytes (+8 so the 8-byte window read is in bounds)
    for (size_t i = 0; i < n; ++i) scalars[i].to_le_bytes(&repr[i * 32]);
    const size_t num_bits = 256;
    const size_t window_size = (size_t)std::ceil(std::log((double)n)) + 2;  // util/msm.rs:268
    const size_t num_buckets = ((size_t)1 << window_size) - 1;
    const size_t num_window = (num_bits + window_size - 1) / window_size;
    std::vector<Bucket> buckets(num_buckets);
    for (size_t idx = num_window; idx-- > 0;) {
        for (size_t k = 0; k < window_size; ++k) *result = result->dbl();
        for (auto& b : buckets) b.kind = Bucket::NONE;
        const size_t skip_bits = idx * window_size, skip_bytes = skip_bits / 8;
        for (size_t i = 0; i < n; ++i) {
            uint8_t v[8] = {0};
            size_t avail = 32 - skip_bytes < 8 ? 32 - skip_bytes : 8;
            memcpy(v, &repr[i * 32 + skip_bytes], avail);
            u64 w; memcpy(&w, v, 8);
            size_t digit = (size_t)(w >> (skip_bits - skip_bytes * 8)) & num_buckets;
            if (digit != 0) buckets[digit - 1].add_assign(bases[i]);
        }
        G1 running = G1::identity();
        for (size_t b = num_buckets; b-- > 0;) {
            running = buckets[b].add(running);
            *result = result->add(running);
        }
    }
}

G1 msm_pippenger(const Fr* scalars, const G1Affine* bases, size_t n, int threads) {
    if (threads <= 1 || n < (size_t)threads) {
        G1 r = G1::identity();
        msm_pippenger_serial(scalars, bases, n, &r);
        return r;
    }
    size_t chunk = (n + threads - 1) / threads;
    std::vector<G1> results(threads, G1::identity());
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) {
        size_t lo = (size_t)t * chunk;
        if (lo >= n) break;
        size_t len = n - lo < chunk ? n - lo : chunk;
        pool.emplace_back([=, &results] { msm_pippenger_serial(scalars + lo, bases + lo, len, &results[t]); });
    }
    for (auto& th : pool) th.join();
    G1 acc = G1::identity();
    for (auto& r : results) acc = acc.add(r);
    return acc;
}

// =====================================================================================================================
// Tower
// =====================================================================================================================
Fq6 Fq6::operator*(const Fq6& o) const {  // Karatsuba over Fq2: 6 Fq2 muls
    Fq2 aa = c0 * o.c0, bb = c1 * o.c1, cc = c2 * o.c2;
    Fq2 t1 = ((c1 + c2) * (o.c1 + o.c2) - bb - cc).mul_xi() + aa;
    Fq2 t2 = (c0 + c1) * (o.c0 + o.c1) - aa - bb + cc.mul_xi();
    Fq2 t3 = (c0 + c2) * (o.c0 + o.c2) - aa - cc + bb;
    return {t1, t2, t3};
}
Fq6 Fq6::mul_by_01(const Fq2& b0, const Fq2& b1) const {
    // (c0 + c1 v + c2 v^2)(b0 + b1 v),  v^3 = xi
    Fq2 a0b0 = c0 * b0, a1b1 = c1 * b1;
    Fq2 r0 = (c2 * b1).mul_xi() + a0b0;
    Fq2 r1 = (c0 + c1) * (b0 + b1) - a0b0 - a1b1;
    Fq2 r2 = c2 * b0 + a1b1;
    return {r0, r1, r2};
}
Fq6 Fq6::inv() const {
    Fq2 t0 = c0.sqr() - (c1 * c2).mul_xi();
    Fq2 t1 = c2.sqr().mul_xi() - c0 * c1;
    Fq2 t2 = c1.sqr() - c0 * c2;
    Fq2 d = (c0 * t0 + (c2 * t1).mul_xi() + (c1 * t2).mul_xi()).inv();
    return {t0 * d, t1 * d, t2 * d};
}
Fq12 Fq12::operator*(const Fq12& o) const {
    Fq6 aa = c0 * o.c0, bb = c1 * o.c1;
    Fq6 m = (c0 + c1) * (o.c0 + o.c1);
    return {aa + bb.mul_v(), m - aa - bb};
}
Fq12 Fq12::sqr() const {
    // (c0 + c1 w)^2 = c0^2 + c1^2 v + 2 c0 c1 w   via  (c0+c1)(c0 + v c1) = c0^2 + v c1^2 + (1+v) c0 c1
    Fq6 ab = c0 * c1;
    Fq6 t = (c0 + c1) * (c0 + c1.mul_v()) - ab - ab.mul_v();
    return {t, ab + ab};
}
Fq12 Fq12::inv() const {
    Fq6 d = (c0.sqr() - c1.sqr().mul_v()).inv();
    return {c0 * d, (c1 * d).neg()};
}
Fq12 Fq12::mul_by_034(const Fq2& l0, const Fq2& l3, const Fq2& l4) const {
    // (a + b w)(A + B w),  A = (l0,0,0), B = (l3,l4,0)
    Fq6 aA = c0.scale(l0);
    Fq6 bB = c1.mul_by_01(l3, l4);
    Fq6 m = (c0 + c1).mul_by_01(l0 + l3, l4);
    return {aA + bB.mul_v(), m - aA - bB};
}

static Fq2 GAMMA[4][6];  // GAMMA[k][i] = xi^(i (p^k - 1)/6), k = 1..3
static G2Affine G2_GEN;
static std::vector<int8_t> ATE_NAF;  // digits of 6u+2, LSB first
static std::once_flag g_init_flag;

static Fq2 fq2_pow(const Fq2& b, const u64 e[4]) {
    Fq2 r = Fq2::one();
    for (int i = 255; i >= 0; --i) {
        r = r.sqr();
        if ((e[i / 64] >> (i % 64)) & 1) r = r * b;
    }
    return r;
}

static Fq fq_from_hex_limbs(u64 l0, u64 l1, u64 l2, u64 l3) {
    u64 t[4] = {l0, l1, l2, l3};
    return Fq::from_raw(t);
}

void init() {
    std::call_once(g_init_flag, [] {
        // gamma1 = xi^((p-1)/6): (p-1)/6 computed by schoolbook division of the 256-bit modulus
        u64 e[4];
        u64 one[4] = {1, 0, 0, 0};
        sub4(e, FQ_MOD.m, one);
        u128 rem = 0;
        for (int i = 3; i >= 0; --i) {
            u128 cur = (rem << 64) | e[i];
            e[i] = (u64)(cur / 6);
            rem = cur % 6;
        }
        Fq2 xi = {Fq::from_u64(9), Fq::from_u64(1)};
        Fq2 g1 = fq2_pow(xi, e);
        Fq2 g2 = g1 * g1.conj();  // xi^((p^2-1)/6), lies in Fq
        Fq2 g3 = g1 * g2;         // xi^((p^3-1)/6)
        Fq2 base[4] = {Fq2::one(), g1, g2, g3};
        for (int k = 1; k <= 3; ++k) {
            GAMMA[k][0] = Fq2::one();
            for (int i = 1; i < 6; ++i) GAMMA[k][i] = GAMMA[k][i - 1] * base[k];
        }
        // G2 generator, canonical integers in LE limbs
        G2_GEN.x.c0 = fq_from_hex_limbs(0x46debd5cd992f6edull, 0x674322d4f75edaddull, 0x426a00665e5c4479ull, 0x1800deef121f1e76ull);
        G2_GEN.x.c1 = fq_from_hex_limbs(0x97e485b7aef312c2ull, 0xf1aa493335a9e712ull, 0x7260bfb731fb5d25ull, 0x198e9393920d483aull);
        G2_GEN.y.c0 = fq_from_hex_limbs(0x4ce6cc0166fa7daaull, 0xe3d1e7690c43d37bull, 0x4aab71808dcb408full, 0x12c85ea5db8c6debull);
        G2_GEN.y.c1 = fq_from_hex_limbs(0x55acdadcd122975bull, 0xbc4b313370b38ef3ull, 0xec9e99ad690c3395ull, 0x090689d0585ff075ull);
        // true NAF of 6u+2 (65 bits)
        u128 n = (u128)6 * BN_U + 2;
        while (n) {
            int8_t d = 0;
            if (n & 1) {
                d = (int8_t)(2 - (int)(n & 3));  // 1 if n mod 4 == 1, -1 if == 3
                if (d == 1) n -= 1; else n += 1;
            }
            ATE_NAF.push_back(d);
            n >>= 1;
        }
    });
}

Fq12 Fq12::frobenius(int power) const {
    // w-power of each tower slot: c0.c0 -> 0, c0.c1 -> 2, c0.c2 -> 4, c1.c0 -> 1, c1.c1 -> 3, c1.c2 -> 5
    const Fq2(&g)[6] = GAMMA[power];
    auto f = [&](const Fq2& a, int i) { return ((power & 1) ? a.conj() : a) * g[i]; };
    return {{f(c0.c0, 0), f(c0.c1, 2), f(c0.c2, 4)}, {f(c1.c0, 1), f(c1.c1, 3), f(c1.c2, 5)}};
}

Fq12 Fq12::cyclotomic_sqr() const {
    // Granger-Scott: view Fq12 = Fq4[w]/(w^3 - t), Fq4 = Fq2[t]/(t^2 - xi), t = w^3;
    // f = g0 + g1 w + g2 w^2 with g0 = (c0.c0, c1.c1), g1 = (c1.c0, c0.c2), g2 = (c0.c1, c1.c2);
    // f^2 = (3 g0^2 - 2 conj(g0)) + (3 t g2^2 + 2 conj(g1)) w + (3 g1^2 - 2 conj(g2)) w^2
    auto fp4_sqr = [](const Fq2& a, const Fq2& b, Fq2& r0, Fq2& r1) {
        Fq2 a2 = a.sqr(), b2 = b.sqr();
        r0 = b2.mul_xi() + a2;
        r1 = (a + b).sqr() - a2 - b2;
    };
    Fq2 a0, a1, b0, b1, d0, d1;
    fp4_sqr(c0.c0, c1.c1, a0, a1);  // g0^2
    fp4_sqr(c1.c0, c0.c2, b0, b1);  // g1^2
    fp4_sqr(c0.c1, c1.c2, d0, d1);  // g2^2
    Fq12 r;
    r.c0.c0 = (a0 - c0.c0).dbl() + a0;
    r.c1.c1 = (a1 + c1.c1).dbl() + a1;
    r.c0.c1 = (b0 - c0.c1).dbl() + b0;
    r.c1.c2 = (b1 + c1.c2).dbl() + b1;
    Fq2 td1 = d1.mul_xi();
    r.c1.c0 = (td1 + c1.c0).dbl() + td1;
    r.c0.c2 = (d0 - c0.c2).dbl() + d0;
    return r;
}

void Fq12::to_le_bytes(uint8_t out[384]) const {
    const Fq* f[12] = {&c0.c0.c0, &c0.c0.c1, &c0.c1.c0, &c0.c1.c1, &c0.c2.c0, &c0.c2.c1,
                       &c1.c0.c0, &c1.c0.c1, &c1.c1.c0, &c1.c1.c1, &c1.c2.c0, &c1.c2.c1};
    for (int i = 0; i < 12; ++i) f[i]->to_le_bytes(out + 32 * i);
}

// =====================================================================================================================
// G2 and the pairing
// =====================================================================================================================
G2Affine G2Affine::generator() { init(); return G2_GEN; }

static Fq2 twist_b() {
    Fq2 xi = {Fq::from_u64(9), Fq::from_u64(1)};
    return Fq2{Fq::from_u64(3), Fq::zero()} * xi.inv();
}
bool G2Affine::is_on_curve() const {
    if (is_identity()) return true;
    return y.sqr() == x.sqr() * x + twist_b();
}

namespace {
struct G2Jac { Fq2 x, y, z; };

// tangent at T (Jacobian over Fq2), line scaled by 2YZ^3 (an Fq2 factor, killed by the final exponentiation):
//   cy = 2YZ * Z^2, cx = -3X^2 Z^2, c0 = 3X^3 - 2Y^2 ;  then T <- 2T (dbl-2009-l)
LineCoeff doubling_step(G2Jac& t) {
    Fq2 a = t.x.sqr(), b = t.y.sqr(), c = b.sqr();
    Fq2 zz = t.z.sqr();
    Fq2 e = a.dbl() + a;  // 3X^2
    Fq2 z3 = (t.y * t.z).dbl();
    LineCoeff l;
    l.cy = z3 * zz;
    l.cx = (e * zz).neg();
    l.c0 = e * t.x - b.dbl();
    Fq2 d = ((t.x + b).sqr() - a - c).dbl();
    Fq2 f = e.sqr();
    Fq2 x3 = f - d.dbl();
    t.y = e * (d - x3) - c.dbl().dbl().dbl();
    t.x = x3;
    t.z = z3;
    return l;
}
// chord through T and affine Q, line scaled by H*Z:  cy = HZ, cx = -R, c0 = R x2 - y2 HZ ;  then T <- T + Q
LineCoeff addition_step(G2Jac& t, const G2Affine& q) {
    Fq2 zz = t.z.sqr();
    Fq2 h = q.x * zz - t.x;
    Fq2 r = q.y * zz * t.z - t.y;
    Fq2 z3 = t.z * h;
    LineCoeff l;
    l.cy = z3;
    l.cx = r.neg();
    l.c0 = r * q.x - q.y * z3;
    Fq2 hh = h.sqr(), hhh = h * hh, v = t.x * hh;
    Fq2 x3 = r.sqr() - hhh - v.dbl();
    t.y = r * (v - x3) - t.y * hhh;
    t.x = x3;
    t.z = z3;
    return l;
}
}  // namespace

G2Prepared::G2Prepared(const G2Affine& q) {
    init();
    infinity = q.is_identity();
    if (infinity) return;
    G2Jac t = {q.x, q.y, Fq2::one()};
    G2Affine nq = q.neg();
    for (int i = (int)ATE_NAF.size() - 2; i >= 0; --i) {
        coeffs.push_back(doubling_step(t));
        if (ATE_NAF[i] == 1) coeffs.push_back(addition_step(t, q));
        else if (ATE_NAF[i] == -1) coeffs.push_back(addition_step(t, nq));
    }
    // Q1 = pi(Q), Q2 = -pi^2(Q)
    G2Affine q1 = {q.x.conj() * GAMMA[1][2], q.y.conj() * GAMMA[1][3]};
    G2Affine q2 = {q.x * GAMMA[2][2], (q.y * GAMMA[2][3]).neg()};
    coeffs.push_back(addition_step(t, q1));
    coeffs.push_back(addition_step(t, q2));
}

static inline void ell(Fq12& f, const LineCoeff& c, const G1Affine& p) {
    f = f.mul_by_034(c.cy.scale(p.y), c.cx.scale(p.x), c.c0);
}

Fq12 multi_miller_loop(const G1Affine* const* ps, const G2Prepared* const* qs, size_t n) {
    init();
    std::vector<size_t> live;
    for (size_t k = 0; k < n; ++k)
        if (!ps[k]->is_identity() && !qs[k]->infinity) live.push_back(k);
    Fq12 f = Fq12::one();
    size_t idx = 0;
    for (int i = (int)ATE_NAF.size() - 2; i >= 0; --i) {
        if (i != (int)ATE_NAF.size() - 2) f = f.sqr();
        for (size_t k : live) ell(f, qs[k]->coeffs[idx], *ps[k]);
        ++idx;
        if (ATE_NAF[i] != 0) {
            for (size_t k : live) ell(f, qs[k]->coeffs[idx], *ps[k]);
            ++idx;
        }
    }
    for (int extra = 0; extra < 2; ++extra) {
        for (size_t k : live) ell(f, qs[k]->coeffs[idx], *ps[k]);
        ++idx;
    }
    return f;
}

static Fq12 exp_by_u(const Fq12& f) {
    Fq12 r = f;
    for (int i = 61; i >= 0; --i) {  // BN_U has 63 bits; bit 62 is the leading one
        r = r.cyclotomic_sqr();
        if ((BN_U >> i) & 1) r = r * f;
    }
    return r;
}

Fq12 final_exponentiation(const Fq12& f_in) {
    init();
    // easy part: f^((p^6 - 1)(p^2 + 1))
    Fq12 f = f_in.conj() * f_in.inv();
    f = f.frobenius(2) * f;
    // hard part (p^4 - p^2 + 1)/r: Devegili-Scott-Dahab vectorial chain y0 y1^2 y2^6 y3^12 y4^18 y5^30 y6^36
    Fq12 fu = exp_by_u(f), fu2 = exp_by_u(fu), fu3 = exp_by_u(fu2);
    Fq12 y0 = f.frobenius(1) * f.frobenius(2) * f.frobenius(3);
    Fq12 y1 = f.conj();
    Fq12 y2 = fu2.frobenius(2);
    Fq12 y3 = fu.frobenius(1).conj();
    Fq12 y4 = (fu * fu2.frobenius(1)).conj();
    Fq12 y5 = fu2.conj();
    Fq12 y6 = (fu3 * fu3.frobenius(1)).conj();
    Fq12 t0 = y6.cyclotomic_sqr() * y4 * y5;
    Fq12 t1 = y3 * y5 * t0;
    t0 = t0 * y2;
    t1 = (t1.cyclotomic_sqr() * t0).cyclotomic_sqr();
    t0 = t1 * y1;
    t1 = t1 * y0;
    return t0.cyclotomic_sqr() * t1;
}

bool kzg_decide(const G1Affine& lhs, const G1Affine& rhs, const G2Affine& g2, const G2Affine& s_g2, Fq12* gt_out) {
    // let terms = [(&lhs, &dk.g2.into()), (&rhs, &(-dk.s_g2).into())];   decider.rs:74
    G2Prepared q0(g2), q1(s_g2.neg());
    const G1Affine* ps[2] = {&lhs, &rhs};
    const G2Prepared* qs[2] = {&q0, &q1};
    Fq12 gt = final_exponentiation(multi_miller_loop(ps, qs, 2));
    if (gt_out) *gt_out = gt;
    return gt.is_one();
}

G2Affine g2_mul_vartime(const G2Affine& q, const u64 k[4]) {
    // Jacobian double-and-add over Fq2 (test-key generation only)
    G2Jac acc = {Fq2::zero(), Fq2::one(), Fq2::zero()};
    bool acc_inf = true;
    for (int i = 255; i >= 0; --i) {
        if (!acc_inf) doubling_step(acc);
        if ((k[i / 64] >> (i % 64)) & 1) {
            if (acc_inf) { acc = {q.x, q.y, Fq2::one()}; acc_inf = false; }
            else addition_step(acc, q);
        }
    }
    if (acc_inf || acc.z.is_zero()) return {Fq2::zero(), Fq2::zero()};
    Fq2 zi = acc.z.inv(), zi2 = zi.sqr();
    return {acc.x * zi2, acc.y * zi2 * zi};
}

}  // namespace oracle

// =====================================================================================================================
// C entry points (ctypes).  Byte formats are those of include/snarkv_cuda.h: scalars 32 B canonical LE,
// G1 64 B x||y canonical LE with (0,0) = identity, G2 128 B x.c0||x.c1||y.c0||y.c1, GT 384 B (12 Fq, tower order).
// Return 0 on success, negative on malformed input.
// =====================================================================================================================
using namespace oracle;

static bool load_g1(const uint8_t* b, G1Affine& p) {
    return Fq::from_le_bytes(b, p.x) && Fq::from_le_bytes(b + 32, p.y) && p.is_on_curve();
}
static void store_g1(const G1Affine& p, uint8_t* b) { p.x.to_le_bytes(b); p.y.to_le_bytes(b + 32); }
static bool load_g2(const uint8_t* b, G2Affine& q) {
    return Fq::from_le_bytes(b, q.x.c0) && Fq::from_le_bytes(b + 32, q.x.c1) && Fq::from_le_bytes(b + 64, q.y.c0) &&
           Fq::from_le_bytes(b + 96, q.y.c1) && q.is_on_curve();
}
static void store_g2(const G2Affine& q, uint8_t* b) {
    q.x.c0.to_le_bytes(b); q.x.c1.to_le_bytes(b + 32); q.y.c0.to_le_bytes(b + 64); q.y.c1.to_le_bytes(b + 96);
}
static int load_terms(const uint8_t* scalars, const uint8_t* points, size_t n, std::vector<Fr>& s, std::vector<G1Affine>& p) {
    s.resize(n); p.resize(n);
    for (size_t i = 0; i < n; ++i) {
        if (!Fr::from_le_bytes(scalars + 32 * i, s[i])) return -2;
        if (!load_g1(points + 64 * i, p[i])) return -3;
    }
    return 0;
}

static inline u64 splitmix64(u64 x) {
    x += 0x9E3779B97F4A7C15ull;
    u64 z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

extern "C" {

int oracle_fp_mul(int field, const uint8_t* a, const uint8_t* b, uint8_t* out) {
    if (field == 0) { Fq x, y; if (!Fq::from_le_bytes(a, x) || !Fq::from_le_bytes(b, y)) return -2; (x * y).to_le_bytes(out); }
    else { Fr x, y; if (!Fr::from_le_bytes(a, x) || !Fr::from_le_bytes(b, y)) return -2; (x * y).to_le_bytes(out); }
    return 0;
}
int oracle_fp_inv(int field, const uint8_t* a, uint8_t* out) {
    if (field == 0) { Fq x; if (!Fq::from_le_bytes(a, x)) return -2; x.inv().to_le_bytes(out); }
    else { Fr x; if (!Fr::from_le_bytes(a, x)) return -2; x.inv().to_le_bytes(out); }
    return 0;
}
// Montgomery form of a canonical value (4 LE u64 limbs) — pins the in-memory layout used by the *_raw C-ABI variants
int oracle_fp_to_mont(int field, const uint8_t* a, uint8_t* out) {
    if (field == 0) { Fq x; if (!Fq::from_le_bytes(a, x)) return -2; memcpy(out, x.v, 32); }
    else { Fr x; if (!Fr::from_le_bytes(a, x)) return -2; memcpy(out, x.v, 32); }
    return 0;
}
int oracle_g1_is_on_curve(const uint8_t* p) { G1Affine a; return load_g1(p, a) ? 1 : 0; }
int oracle_g1_add(const uint8_t* a, const uint8_t* b, uint8_t* out) {
    G1Affine x, y; if (!load_g1(a, x) || !load_g1(b, y)) return -3;
    store_g1(G1::from_affine(x).add_mixed(y).to_affine(), out); return 0;
}
int oracle_g1_mul(const uint8_t* p, const uint8_t* scalar, uint8_t* out) {
    G1Affine a; Fr s; if (!load_g1(p, a)) return -3; if (!Fr::from_le_bytes(scalar, s)) return -2;
    store_g1(g1_mul_ct(a, s).to_affine(), out); return 0;
}
int oracle_g2_generator(uint8_t* out) { store_g2(G2Affine::generator(), out); return 0; }
int oracle_g2_mul(const uint8_t* q, const uint8_t* scalar, uint8_t* out) {
    init();
    G2Affine a; if (!load_g2(q, a)) return -3;
    u64 k[4]; memcpy(k, scalar, 32);
    store_g2(g2_mul_vartime(a, k), out); return 0;
}
int oracle_g2_is_on_curve(const uint8_t* q) { G2Affine a; return load_g2(q, a) ? 1 : 0; }

// loader/native.rs:61-71
int oracle_msm_native(const uint8_t* scalars, const uint8_t* points, size_t n, uint8_t* out) {
    if (n == 0) return -1;  // .unwrap() panics on an empty slice (native.rs:69)
    std::vector<Fr> s; std::vector<G1Affine> p;
    int rc = load_terms(scalars, points, n, s, p); if (rc) return rc;
    store_g1(msm_native_fold(s.data(), p.data(), n), out); return 0;
}
// util/msm.rs:308-343 (+ caller's .to_affine()); threads <= 1 = the non-`parallel` build
int oracle_msm_pippenger(const uint8_t* scalars, const uint8_t* points, size_t n, int threads, uint8_t* out) {
    if (n == 0) return -1;
    std::vector<Fr> s; std::vector<G1Affine> p;
    int rc = load_terms(scalars, points, n, s, p); if (rc) return rc;
    store_g1(msm_pippenger(s.data(), p.data(), n, threads).to_affine(), out); return 0;
}
// The same call on halo2curves' in-memory layout (Montgomery limbs, what `&[Fr]` / `&[G1Affine]` are in the Rust reference):
// no parsing, no validation — only the algorithm of util/msm.rs:308-343 is inside the call.  This is the entry bench.py times.
int oracle_msm_pippenger_raw(const uint8_t* scalars_mont, const uint8_t* points_mont, size_t n, int threads, uint8_t* out) {
    if (n == 0) return -1;
    static_assert(sizeof(Fr) == 32 && sizeof(G1Affine) == 64, "layout");
    const Fr* s = reinterpret_cast<const Fr*>(scalars_mont);
    const G1Affine* p = reinterpret_cast<const G1Affine*>(points_mont);
    store_g1(msm_pippenger(s, p, n, threads).to_affine(), out); return 0;
}
// canonical -> Montgomery layout for n field elements of `field` (0 = Fq, 1 = Fr), multi-threaded (bench input preparation)
int oracle_to_mont_batch(int field, const uint8_t* in, size_t n, int threads, uint8_t* out) {
    if (threads < 1) threads = 1;
    std::vector<std::thread> pool;
    std::vector<int> rc(threads, 0);
    size_t chunk = (n + threads - 1) / threads;
    for (int t = 0; t < threads; ++t) {
        size_t lo = (size_t)t * chunk, hi = lo + chunk < n ? lo + chunk : n;
        if (lo >= hi) break;
        pool.emplace_back([=, &rc] {
            for (size_t i = lo; i < hi; ++i) {
                if (field == 0) { Fq x; if (!Fq::from_le_bytes(in + 32 * i, x)) { rc[t] = -2; return; } memcpy(out + 32 * i, x.v, 32); }
                else { Fr x; if (!Fr::from_le_bytes(in + 32 * i, x)) { rc[t] = -2; return; } memcpy(out + 32 * i, x.v, 32); }
            }
        });
    }
    for (auto& th : pool) th.join();
    for (int r : rc) if (r) return r;
    return 0;
}
// pcs/kzg/decider.rs:70-82;  *accept = 1/0;  gt_out may be NULL
int oracle_kzg_decide(const uint8_t* lhs, const uint8_t* rhs, const uint8_t* g2, const uint8_t* s_g2, uint8_t* accept, uint8_t* gt_out) {
    G1Affine l, r; G2Affine a, b;
    if (!load_g1(lhs, l) || !load_g1(rhs, r)) return -3;
    if (!load_g2(g2, a) || !load_g2(s_g2, b)) return -4;
    Fq12 gt; *accept = kzg_decide(l, r, a, b, &gt) ? 1 : 0;
    if (gt_out) gt.to_le_bytes(gt_out);
    return 0;
}
// pcs/kzg/decider.rs:84-93 over N accumulators, chunked over `threads`; accept[i] per accumulator (the reference
// aborts at the first failure; per-item flags are a superset of that information).  `hoist` != 0 prepares the two G2
// points once instead of once per accumulator (BASELINE.md B3').
int oracle_kzg_decide_batch(const uint8_t* lhs, const uint8_t* rhs, size_t n, const uint8_t* g2, const uint8_t* s_g2,
                            int threads, int hoist, uint8_t* accept, uint8_t* gt_out) {
    G2Affine a, b;
    if (!load_g2(g2, a) || !load_g2(s_g2, b)) return -4;
    std::vector<G1Affine> l(n), r(n);
    for (size_t i = 0; i < n; ++i)
        if (!load_g1(lhs + 64 * i, l[i]) || !load_g1(rhs + 64 * i, r[i])) return -3;
    G2Prepared q0(a), q1(b.neg());
    if (threads < 1) threads = 1;
    auto work = [&](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; ++i) {
            Fq12 gt; bool ok;
            if (hoist) {
                const G1Affine* ps[2] = {&l[i], &r[i]};
                const G2Prepared* qs[2] = {&q0, &q1};
                gt = final_exponentiation(multi_miller_loop(ps, qs, 2));
                ok = gt.is_one();
            } else ok = kzg_decide(l[i], r[i], a, b, &gt);
            accept[i] = ok ? 1 : 0;
            if (gt_out) gt.to_le_bytes(gt_out + 384 * i);
        }
    };
    std::vector<std::thread> pool;
    size_t chunk = (n + threads - 1) / threads;
    for (int t = 0; t < threads; ++t) {
        size_t lo = (size_t)t * chunk, hi = lo + chunk < n ? lo + chunk : n;
        if (lo >= hi) break;
        pool.emplace_back(work, lo, hi);
    }
    for (auto& th : pool) th.join();
    return 0;
}
// pcs/kzg/accumulation.rs:41-63: powers_of_r = r.powers(n) (loader.rs:71-78), lhs = sum r^i lhs_i, rhs likewise,
// each evaluated by NativeLoader::multi_scalar_multiplication
int oracle_kzg_accumulate(const uint8_t* lhs, const uint8_t* rhs, size_t n, const uint8_t* r, uint8_t* out_lhs, uint8_t* out_rhs) {
    if (n == 0) return -1;
    Fr rr; if (!Fr::from_le_bytes(r, rr)) return -2;
    std::vector<Fr> pw(n); pw[0] = Fr::one();
    for (size_t i = 1; i < n; ++i) pw[i] = pw[i - 1] * rr;
    std::vector<G1Affine> l(n), q(n);
    for (size_t i = 0; i < n; ++i)
        if (!load_g1(lhs + 64 * i, l[i]) || !load_g1(rhs + 64 * i, q[i])) return -3;
    store_g1(msm_native_fold(pw.data(), l.data(), n), out_lhs);
    store_g1(msm_native_fold(pw.data(), q.data(), n), out_rhs);
    return 0;
}

// ---- synthetic workload, same definition as bn254_model.synth_* and the CUDA generator kernel -------------------------
void oracle_synth_scalars(u64 seed, u64 start, size_t n, uint8_t* out) {
    for (size_t i = 0; i < n; ++i) {
        u64 l[4];
        for (int k = 0; k < 4; ++k) l[k] = splitmix64(seed * 0x100000001B3ull + (start + i) * 4 + k);
        l[3] &= (1ull << 62) - 1;
        if (geq(l, FR_MOD.m)) sub4(l, l, FR_MOD.m);
        memcpy(out + 32 * i, l, 32);
    }
}
void oracle_synth_point_scalars(u64 seed, u64 start, size_t n, u64* out) {
    for (size_t i = 0; i < n; ++i) out[i] = splitmix64(seed * 0x100000001B3ull + 0x5151515151515151ull + start + i) | 1;
}
// P_i = [t_i] G, affine canonical bytes.  Same definition as before; computed with a fixed-base comb (8 windows of 8 bits over
// the 64-bit t_i: table[w][d] = d 2^(8w) G, 7 mixed additions per point) so that the full 2^24-term workload of the bench's
// CPU arm is generated in seconds.  oracle_synth_points_slow keeps the literal double-and-add for the cross-check in tests/.
static const G1Affine* synth_comb_table() {
    static std::vector<G1Affine> table;
    static std::once_flag once;
    std::call_once(once, [] {
        std::vector<G1> jac(8 * 256);
        G1 base = G1::from_affine(G1Affine::generator());
        for (int w = 0; w < 8; ++w) {
            G1 acc = G1::identity();
            for (int d = 0; d < 256; ++d) {
                jac[w * 256 + d] = acc;
                acc = acc.add(base);
            }
            base = acc;   // 256 * previous base
        }
        table.resize(8 * 256);
        g1_batch_to_affine(jac.data(), table.data(), jac.size());
    });
    return table.data();
}
void oracle_synth_points(u64 seed, u64 start, size_t n, int threads, uint8_t* out) {
    if (threads < 1) threads = 1;
    const G1Affine* table = synth_comb_table();
    auto work = [&](size_t lo, size_t hi) {
        const size_t B = 1024;
        std::vector<G1> jac(B); std::vector<G1Affine> aff(B);
        for (size_t base = lo; base < hi; base += B) {
            size_t m = hi - base < B ? hi - base : B;
            for (size_t j = 0; j < m; ++j) {
                const u64 t = splitmix64(seed * 0x100000001B3ull + 0x5151515151515151ull + start + base + j) | 1;
                G1 acc = G1::from_affine(table[t & 0xff]);      // t is odd: the low digit is never 0
                for (int w = 1; w < 8; ++w) {
                    const unsigned d = (unsigned)(t >> (8 * w)) & 0xffu;
                    if (d) acc = acc.add_mixed(table[w * 256 + d]);
                }
                jac[j] = acc;
            }
            g1_batch_to_affine(jac.data(), aff.data(), m);
            for (size_t j = 0; j < m; ++j) store_g1(aff[j], out + 64 * (base + j));
        }
    };
    std::vector<std::thread> pool;
    size_t chunk = (n + threads - 1) / threads;
    for (int t = 0; t < threads; ++t) {
        size_t lo = (size_t)t * chunk, hi = lo + chunk < n ? lo + chunk : n;
        if (lo >= hi) break;
        pool.emplace_back(work, lo, hi);
    }
    for (auto& th : pool) th.join();
}
void oracle_synth_points_slow(u64 seed, u64 start, size_t n, uint8_t* out) {
    const G1Affine g = G1Affine::generator();
    for (size_t j = 0; j < n; ++j) {
        u64 t[4] = {splitmix64(seed * 0x100000001B3ull + 0x5151515151515151ull + start + j) | 1, 0, 0, 0};
        store_g1(g1_mul_vartime(g, t).to_affine(), out + 64 * j);
    }
}
// checksum for full-size runs: (sum_i s_i * t_i mod r) * G, where P_i = [t_i] G
int oracle_msm_expected_from_dlogs(const uint8_t* scalars, const u64* t, size_t n, uint8_t* out) {
    Fr acc = Fr::zero();
    for (size_t i = 0; i < n; ++i) {
        Fr s; if (!Fr::from_le_bytes(scalars + 32 * i, s)) return -2;
        acc = acc + s * Fr::from_u64(t[i]);
    }
    store_g1(g1_mul_ct(G1Affine::generator(), acc).to_affine(), out);
    return 0;
}
u64 oracle_mulmod_count_reset() { u64 c = g_mulmod_count; g_mulmod_count = 0; return c; }

}  // extern "C"
