// CPU-only driver for snark_verifier_b200/host/plonk_eval.hpp (run by tests/test_plonk_eval.py): prints the straight-line program the
// C++ compiler emits for the StandardPlonk-shaped protocol so that it can be compared, instruction by instruction, with the
// Python compiler's.  usage: plonk_compile_test k num_instance
#include <cstdio>
#include <cstdlib>
#include <string>

#include "../snark_verifier_b200/host/plonk_eval.hpp"

static void print_program(const snarkv::FrProgram& prog) {
    printf("n_regs %u n_inputs %zu\n", prog.n_regs, prog.n_inputs);
    for (const auto& in : prog.instrs) printf("i %u %u %u %u\n", in.op, in.dst, in.a, in.b);
    for (const auto& c : prog.consts) {
        printf("c ");
        for (int i = 31; i >= 0; --i) printf("%02x", c[i]);
        printf("\n");
    }
    printf("o");
    for (uint32_t o : prog.outputs) printf(" %u", o);
    printf("\n");
}

// usage: plonk_compile_test k num_instance                     quotient evaluation of the StandardPlonk-shaped protocol
//        plonk_compile_test gwc19|bdfg21 queries.bin            MSM-scalar program of the multi-open verifier; queries.bin =
//                                                               u32 num_polys | u32 nq | nq x (u32 poly | 32-byte shift)
int main(int argc, char** argv) {
    if (argc < 3) return 2;
    using namespace snarkv;
    const std::string mode = argv[1];
    if (mode == "gwc19" || mode == "bdfg21") {
        FILE* f = fopen(argv[2], "rb");
        if (!f) return 2;
        uint32_t npoly, nq;
        if (fread(&npoly, 4, 1, f) != 1 || fread(&nq, 4, 1, f) != 1) return 2;
        std::vector<plonk::ShiftQuery> qs(nq);
        for (auto& q : qs) {
            uint32_t p32;
            if (fread(&p32, 4, 1, f) != 1 || fread(q.shift.v, 1, 32, f) != 32) return 2;
            q.poly = p32;
        }
        fclose(f);
        const plonk::MsmScalarProgram mp = mode == "gwc19" ? plonk::compile_gwc19_msm_scalars(qs, npoly) : plonk::compile_bdfg21_msm_scalars(qs, npoly);
        print_program(mp.program);
        printf("l");
        for (const auto& sl : mp.lhs_slots) printf(" %c%u", sl.kind, sl.idx);
        printf("\nr");
        for (const auto& sl : mp.rhs_slots) printf(" %c%u", sl.kind, sl.idx);
        printf("\n");
        return 0;
    }
    const plonk::QuotientProtocol proto = plonk::standard_plonk_like_protocol(atoi(argv[1]), (size_t)atoi(argv[2]));
    print_program(plonk::compile_quotient_evaluation(proto));
    return 0;
}
