// CPU-only driver for the C++ multi-open verifiers (host/cuda_loader.hpp `BasicGwc19`, host/pcs.hpp `BasicBdfg21`), run by
// tests/test_pcs_mirror.py: the verifiers run over a RECORDING loader whose multi_scalar_multiplication prints its (scalar, base)
// pairs, and the Python test compares them with what the Python mirror hands to its loader.
// Input (binary): u8 scheme (0 = GWC19, 1 = SHPLONK) | u32 npoly | u32 nq | z | nq x (u32 poly | shift | eval) | npoly x commitment |
//                 svk g | GWC19: v | u | u32 nw | nw x W      SHPLONK: mu | gamma | W | z' | W'
#include <cstdio>
#include <cstdlib>

#include "../snark_verifier_b200/host/pcs.hpp"
#include "../snark_verifier_b200/host/plonk_eval.hpp"

using namespace snarkv;

struct RecordingLoader {
    G1Affine multi_scalar_multiplication(const std::vector<std::pair<const Fr*, const G1Affine*>>& pairs) {
        printf("msm %zu\n", pairs.size());
        for (const auto& pr : pairs) {
            for (int i = 31; i >= 0; --i) printf("%02x", (*pr.first)[i]);
            printf(" ");
            for (int i = 0; i < 64; ++i) printf("%02x", (*pr.second)[i]);
            printf("\n");
        }
        return G1Affine{};
    }
};

// FrOps over the setup-time field arithmetic of plonk_eval.hpp
static plonk::Fe fe_of(const Fr& a) { plonk::Fe r; memcpy(r.v, a.data(), 32); return r; }
static Fr op_add(const Fr& a, const Fr& b) { return plonk::fe_to_bytes(plonk::fe_add(fe_of(a), fe_of(b))); }
static Fr op_mul(const Fr& a, const Fr& b) { return plonk::fe_to_bytes(plonk::fe_mul(fe_of(a), fe_of(b))); }
static Fr op_neg(const Fr& a) { return plonk::fe_to_bytes(plonk::fe_neg(fe_of(a))); }
static Fr op_inv(const Fr& a) { const plonk::Fe x = fe_of(a); return x.is_zero() ? a : plonk::fe_to_bytes(plonk::fe_inv(x)); }

template <size_t N> static std::array<uint8_t, N> rd(FILE* f) {
    std::array<uint8_t, N> a;
    if (fread(a.data(), 1, N, f) != N) { fprintf(stderr, "short read\n"); exit(2); }
    return a;
}
static uint32_t rd32(FILE* f) { uint32_t v; if (fread(&v, 4, 1, f) != 1) exit(2); return v; }

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 2;
    uint8_t scheme;
    if (fread(&scheme, 1, 1, f) != 1) return 2;
    const uint32_t npoly = rd32(f), nq = rd32(f);
    const Fr z = rd<32>(f);
    std::vector<Query> queries(nq);
    for (auto& q : queries) { q.poly = rd32(f); q.shift = rd<32>(f); q.eval = rd<32>(f); }
    std::vector<G1Affine> commits(npoly);
    for (auto& c : commits) c = rd<64>(f);
    const G1Affine g = rd<64>(f);
    Fr one{}; one[0] = 1;
    const FrOps ops{op_add, op_mul, one, op_neg, op_inv};
    RecordingLoader loader;
    std::vector<BasicMsm<RecordingLoader>> cm;
    for (const auto& c : commits) cm.push_back(BasicMsm<RecordingLoader>::base(loader, ops, c));
    if (scheme == 0) {
        Gwc19Proof proof;
        proof.v = rd<32>(f); proof.u = rd<32>(f);
        const uint32_t nw = rd32(f);
        proof.ws.resize(nw);
        for (auto& w : proof.ws) w = rd<64>(f);
        BasicGwc19<RecordingLoader>::verify(loader, ops, g, cm, z, queries, proof);
    } else {
        Bdfg21Proof proof;
        proof.mu = rd<32>(f); proof.gamma = rd<32>(f); proof.w = rd<64>(f); proof.z_prime = rd<32>(f); proof.w_prime = rd<64>(f);
        BasicBdfg21<RecordingLoader>::verify(loader, ops, g, cm, z, queries, proof);
    }
    fclose(f);
    return 0;
}
