// CPU-only driver for snark_verifier_b200/host/plonk_eval.hpp (run by tests/test_plonk_eval.py): prints the straight-line program the
// C++ compiler emits for the StandardPlonk-shaped protocol so that it can be compared, instruction by instruction, with the
// Python compiler's.  usage: plonk_compile_test k num_instance
#include <cstdio>
#include <cstdlib>

#include "../snark_verifier_b200/host/plonk_eval.hpp"

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    using namespace snarkv;
    const plonk::QuotientProtocol proto = plonk::standard_plonk_like_protocol(atoi(argv[1]), (size_t)atoi(argv[2]));
    const FrProgram prog = plonk::compile_quotient_evaluation(proto);
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
    return 0;
}
