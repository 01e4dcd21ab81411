// C++ host-mirror test (run by tests/test_gpu_parity.py::test_cpp_host_mirror on the GPU box): drives
// snark_verifier_b200/host/cuda_loader.hpp through the same golden inputs as the Python suite.  Input file format (binary):
//   u32 n | n x 32 B scalars | n x 64 B points | 64 B expected MSM | 128 B g2 | 128 B s_g2 | 64 B lhs_ok | 64 B rhs_ok | 64 B rhs_bad
//   | GWC19 section (see main)
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../snark_verifier_b200/host/cuda_loader.hpp"
#include "../snark_verifier_b200/host/plonk_eval.hpp"

using namespace snarkv;

template <size_t N> static std::array<uint8_t, N> rd(FILE* f) {
    std::array<uint8_t, N> a;
    if (fread(a.data(), 1, N, f) != N) { fprintf(stderr, "short read\n"); exit(2); }
    return a;
}

// ---- minimal Fr arithmetic for the Msm bookkeeping (FrOps policy): 4 x u64 limbs, add / double-and-add multiplication -------
static const uint64_t R_MOD[4] = {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
struct U256 { uint64_t v[4]; };
static U256 from_fr(const Fr& a) { U256 r; memcpy(r.v, a.data(), 32); return r; }
static Fr to_fr(const U256& a) { Fr r; memcpy(r.data(), a.v, 32); return r; }
static bool geq_mod(const U256& a) {
    for (int i = 3; i >= 0; --i) { if (a.v[i] > R_MOD[i]) return true; if (a.v[i] < R_MOD[i]) return false; }
    return true;
}
static U256 add_mod(const U256& a, const U256& b) {   // a, b < r < 2^254: no overflow out of 256 bits
    U256 r; unsigned __int128 c = 0;
    for (int i = 0; i < 4; ++i) { c += (unsigned __int128)a.v[i] + b.v[i]; r.v[i] = (uint64_t)c; c >>= 64; }
    if (geq_mod(r)) { unsigned __int128 br = 0; for (int i = 0; i < 4; ++i) { unsigned __int128 d = (unsigned __int128)r.v[i] - R_MOD[i] - br; r.v[i] = (uint64_t)d; br = (d >> 64) & 1; } }
    return r;
}
static Fr fr_add(const Fr& a, const Fr& b) { return to_fr(add_mod(from_fr(a), from_fr(b))); }
static Fr fr_mul(const Fr& a, const Fr& b) {
    U256 x = from_fr(a), y = from_fr(b), acc{};
    for (int bit = 255; bit >= 0; --bit) {
        acc = add_mod(acc, acc);
        if ((y.v[bit >> 6] >> (bit & 63)) & 1) acc = add_mod(acc, x);
    }
    return to_fr(acc);
}
static Fr fr_neg(const Fr& a) {
    U256 x = from_fr(a), r; bool zero = !(x.v[0] | x.v[1] | x.v[2] | x.v[3]);
    if (zero) return a;
    unsigned __int128 br = 0;
    for (int i = 0; i < 4; ++i) { unsigned __int128 d = (unsigned __int128)R_MOD[i] - x.v[i] - br; r.v[i] = (uint64_t)d; br = (d >> 64) & 1; }
    return to_fr(r);
}

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 2;
    uint32_t n;
    if (fread(&n, 4, 1, f) != 1) return 2;
    std::vector<Fr> s(n); std::vector<G1Affine> p(n);
    for (auto& x : s) x = rd<32>(f);
    for (auto& x : p) x = rd<64>(f);
    G1Affine expected = rd<64>(f);
    KzgDecidingKey dk;
    dk.g2 = rd<128>(f); dk.s_g2 = rd<128>(f);
    dk.g = G1Affine{};
    G1Affine lhs_ok = rd<64>(f), rhs_ok = rd<64>(f), rhs_bad = rd<64>(f);
    // GWC19 section: u32 npoly | u32 nq | z | v | u | nq x (u32 poly | shift | eval) | npoly x commitment | u32 nw | nw x W |
    //                g2 | s_g2 | svk g | expected lhs | expected rhs
    uint32_t npoly, nq, nw;
    if (fread(&npoly, 4, 1, f) != 1 || fread(&nq, 4, 1, f) != 1) return 2;
    Fr gz = rd<32>(f), gv = rd<32>(f), gu = rd<32>(f);
    std::vector<Query> queries(nq);
    for (auto& q : queries) { uint32_t p32; if (fread(&p32, 4, 1, f) != 1) return 2; q.poly = p32; q.shift = rd<32>(f); q.eval = rd<32>(f); }
    std::vector<G1Affine> commits(npoly);
    for (auto& c : commits) c = rd<64>(f);
    if (fread(&nw, 4, 1, f) != 1) return 2;
    Gwc19Proof gproof{gv, std::vector<G1Affine>(nw), gu};
    for (auto& w : gproof.ws) w = rd<64>(f);
    KzgDecidingKey gdk;
    gdk.g2 = rd<128>(f); gdk.s_g2 = rd<128>(f); gdk.g = rd<64>(f);
    G1Affine exp_lhs = rd<64>(f), exp_rhs = rd<64>(f);
    // scalar-evaluation section (tests/golden/plonk_eval.json, case k = 8, 3 instances): u32 rows | u32 n_in | u32 n_out | inputs | expected
    uint32_t pe_rows, pe_in, pe_out;
    if (fread(&pe_rows, 4, 1, f) != 1 || fread(&pe_in, 4, 1, f) != 1 || fread(&pe_out, 4, 1, f) != 1) return 2;
    std::vector<Fr> pe_inputs(pe_rows * pe_in), pe_expected(pe_rows * pe_out);
    for (auto& x : pe_inputs) x = rd<32>(f);
    for (auto& x : pe_expected) x = rd<32>(f);
    fclose(f);

    CudaLoader loader(0);
    std::vector<std::pair<const Fr*, const G1Affine*>> pairs;
    for (uint32_t i = 0; i < n; ++i) pairs.emplace_back(&s[i], &p[i]);
    if (loader.multi_scalar_multiplication(pairs) != expected) { fprintf(stderr, "MSM mismatch\n"); return 1; }
    bool threw = false;
    try { loader.multi_scalar_multiplication({}); } catch (const Error&) { threw = true; }   // native.rs:69 panics on empty
    if (!threw) { fprintf(stderr, "empty MSM did not fail\n"); return 1; }
    try { loader.ec_point_assert_eq("points differ", lhs_ok, rhs_ok); threw = false; } catch (const AssertionFailure&) { threw = true; }
    if (!threw) return 1;

    KzgAs as(loader, dk);
    as.decide({lhs_ok, rhs_ok});
    as.decide_all({{lhs_ok, rhs_ok}, {lhs_ok, rhs_ok}});
    as.decide_all({});
    threw = false;
    try { as.decide_all({{lhs_ok, rhs_ok}, {lhs_ok, rhs_bad}}); } catch (const AssertionFailure& e) {
        threw = std::string(e.what()) == KzgAs::ASSERTION;
    }
    if (!threw) { fprintf(stderr, "tampered accumulator was not rejected with the reference's message\n"); return 1; }
    // accumulate two copies of a valid accumulator with r = 5: still valid
    Fr r{}; r[0] = 5;
    KzgAccumulator acc = as.verify({{lhs_ok, rhs_ok}, {lhs_ok, rhs_ok}}, r);
    as.decide(acc);
    // LimbsEncoding<4, 68>::from_repr (accumulator.rs:57-81): split the valid accumulator into 16 limbs and rebuild it
    {
        std::vector<Fr> limbs;
        for (const G1Affine* pt : {&lhs_ok, &rhs_ok})
            for (int c = 0; c < 2; ++c)
                for (int i = 0; i < 4; ++i) {   // limb i = bits [68 i, 68 i + 68) of the coordinate
                    Fr l{};
                    for (int b = 0; b < 68; ++b) {
                        const int src = 68 * i + b;
                        if (src < 256 && (((*pt)[32 * c + src / 8] >> (src % 8)) & 1)) l[b / 8] |= (uint8_t)(1u << (b % 8));
                    }
                    limbs.push_back(l);
                }
        KzgAccumulator back = LimbsEncoding<4, 68>::from_repr(loader, limbs);
        if (back.lhs != lhs_ok || back.rhs != rhs_ok) { fprintf(stderr, "LimbsEncoding round trip failed\n"); return 1; }
        as.decide(back);
        limbs[0][0] ^= 1;   // off the curve: the reference panics, the mirror throws
        threw = false;
        try { LimbsEncoding<4, 68>::from_repr(loader, limbs); } catch (const Error&) { threw = true; }
        if (!threw) { fprintf(stderr, "off-curve limbs were accepted\n"); return 1; }
    }
    // FrProgram: out = (in0 * in0 + 5) ^ -1 for three proofs; 1 / (x^2 + 5) * (x^2 + 5) must be 1
    {
        FrProgram prog;
        Fr five{}; five[0] = 5;
        prog.consts = {five};
        prog.n_inputs = 1;
        prog.n_regs = 4;
        prog.instrs = {{SNARKV_FR_OP_INPUT, 0, 0, 0}, {SNARKV_FR_OP_MUL, 1, 0, 0}, {SNARKV_FR_OP_CONST, 2, 0, 0}, {SNARKV_FR_OP_ADD, 1, 1, 2},
                       {SNARKV_FR_OP_INV, 3, 1, 0}, {SNARKV_FR_OP_MUL, 3, 3, 1}};
        prog.outputs = {3, 1};
        std::vector<Fr> in(3);
        in[0][0] = 2; in[1][0] = 3; in[2][5] = 9;
        std::vector<Fr> out = prog.eval_batch(loader, in, 3);
        Fr one{}; one[0] = 1;
        Fr nine{}; nine[0] = 9;
        if (out[0] != one || out[2] != one || out[4] != one || out[1] != nine) { fprintf(stderr, "FrProgram mismatch\n"); return 1; }
        prog.instrs[1].a = 3;   // reads a register before it is written
        threw = false;
        try { prog.eval_batch(loader, in, 3); } catch (const Error&) { threw = true; }
        if (!threw) { fprintf(stderr, "malformed program was accepted\n"); return 1; }
    }
    // Gwc19::verify (gwc19.rs:45-82) through the C++ Msm algebra and the device MSM: same accumulator bytes as the Python mirror over
    // the NativeLoader fold, and it decides under the fixture's key
    {
        Fr one{}; one[0] = 1;
        FrOps ops{fr_add, fr_mul, one, fr_neg};
        std::vector<Msm> cm;
        for (const auto& c : commits) cm.push_back(Msm::base(loader, ops, c));
        KzgAccumulator acc = Gwc19::verify(loader, ops, gdk.g, cm, gz, queries, gproof);
        if (acc.lhs != exp_lhs || acc.rhs != exp_rhs) { fprintf(stderr, "Gwc19::verify accumulator mismatch\n"); return 1; }
        KzgAs gas(loader, gdk);
        gas.decide(acc);
        queries[1].eval[0] ^= 1;   // tampered evaluation
        threw = false;
        try { gas.decide(Gwc19::verify(loader, ops, gdk.g, cm, gz, queries, gproof)); } catch (const AssertionFailure&) { threw = true; }
        if (!threw) { fprintf(stderr, "tampered GWC19 proof was accepted\n"); return 1; }
    }
    // the C++ compiler (host/plonk_eval.hpp: Expression::evaluate over a ProgramBuilder) + the device program, against the golden rows
    {
        const plonk::QuotientProtocol proto = plonk::standard_plonk_like_protocol(8, 3);
        const FrProgram prog = plonk::compile_quotient_evaluation(proto);
        if (prog.n_inputs != pe_in || prog.outputs.size() != pe_out) { fprintf(stderr, "program shape mismatch\n"); return 1; }
        if (prog.eval_batch(loader, pe_inputs, pe_rows) != pe_expected) { fprintf(stderr, "quotient evaluation mismatch\n"); return 1; }
    }
    printf("host mirror ok\n");
    return 0;
}
