// C++ host-mirror test (run by tests/test_gpu_parity.py::test_cpp_host_mirror on the GPU box): drives
// snark_verifier_b200/host/cuda_loader.hpp through the same golden inputs as the Python suite.  Input file format (binary):
//   u32 n | n x 32 B scalars | n x 64 B points | 64 B expected MSM | 128 B g2 | 128 B s_g2 | 64 B lhs_ok | 64 B rhs_ok | 64 B rhs_bad
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../snark_verifier_b200/host/cuda_loader.hpp"

using namespace snarkv;

template <size_t N> static std::array<uint8_t, N> rd(FILE* f) {
    std::array<uint8_t, N> a;
    if (fread(a.data(), 1, N, f) != N) { fprintf(stderr, "short read\n"); exit(2); }
    return a;
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
    printf("host mirror ok\n");
    return 0;
}
