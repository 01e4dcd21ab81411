// fr_program.cu — one straight-line Fr program evaluated for a batch of proofs (the third "next" row).
//
// Replaces, for m proofs of ONE protocol, the per-proof scalar work the native verifier does before it builds its MSMs:
//   snark-verifier/src/verifier/plonk/protocol.rs:211-283   CommonPolynomialEvaluation::{new, evaluate}  (z^n, 1/(z^n - 1), L_i(z))
//   snark-verifier/src/verifier/plonk/protocol.rs:333-392   Expression::evaluate                        (the quotient numerator)
//   snark-verifier/src/verifier/plonk/proof.rs:306-349      PlonkProof::evaluations                     (instance evaluations)
//   snark-verifier/src/verifier/plonk/proof.rs:298-303      numerator * zn_minus_one_inv                (quotient evaluation)
// The protocol is the same for every proof of a batch, so the expression tree is the same: the host flattens it once into a
// register program (snark_verifier_b200/plonk_eval.py) and one thread per proof runs it.  The register file lives in global
// memory as reg[r][proof] (32-byte values, Montgomery form): consecutive threads touch consecutive 32-byte words, and at
// m = 4096 one register row is 128 KB, so the whole file stays in L2.  The program itself is read through the uniform
// read-only path (every thread of a warp reads the same instruction).
//
// Instruction set (struct snarkv_fr_instr {op, dst, a, b}):
//   INPUT  dst <- inputs[proof][a]          CONST  dst <- consts[a]
//   ADD    dst <- reg[a] + reg[b]           SUB    dst <- reg[a] - reg[b]         MUL  dst <- reg[a] * reg[b]
//   NEG    dst <- -reg[a]                   INV    dst <- 1 / reg[a], and 0 when reg[a] = 0
//   NZ     dst <- reg[a] == 0 ? 1 : reg[a]  KEEPZ  dst <- reg[b] == 0 ? 0 : reg[a]   (batch_invert's zero skipping as selects)
// INV follows `ScalarLoader::batch_invert` (loader.rs:255-262 / util/arithmetic.rs:47-74): a zero is left untouched.
// pow_const (loader.rs:52-69) needs no opcode: the host emits its exact square-and-multiply sequence as MULs.
#include "ctx.hpp"
#include "fp.cuh"

namespace snarkv {

// plain (coherent) loads: the register file is written by this very kernel, so the read-only path (__ldg) is not allowed
__device__ __forceinline__ Fr fr_load_rw(const void* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    const uint4 lo = q[0], hi = q[1];
    Fr r;
    r.v[0] = lo.x; r.v[1] = lo.y; r.v[2] = lo.z; r.v[3] = lo.w;
    r.v[4] = hi.x; r.v[5] = hi.y; r.v[6] = hi.z; r.v[7] = hi.w;
    return r;
}

__global__ void __launch_bounds__(128) k_fr_program(const uint4* __restrict__ prog, uint32_t n_instr, const uint8_t* __restrict__ consts,
                                                    const uint8_t* __restrict__ inputs, uint32_t n_inputs, size_t m, int format,
                                                    uint8_t* regs, const uint32_t* __restrict__ out_regs, uint32_t n_out,
                                                    uint8_t* __restrict__ outputs) {
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= m) return;
    auto reg = [&](uint32_t r) -> uint8_t* { return regs + ((size_t)r * m + p) * 32; };
#pragma unroll 1
    for (uint32_t pc = 0; pc < n_instr; ++pc) {
        const uint4 ins = __ldg(&prog[pc]);
        Fr v;
        switch (ins.x) {
            case SNARKV_FR_OP_INPUT:
                v = fp_load<FR>(inputs + (p * n_inputs + ins.z) * 32);
                if (format == SNARKV_CANONICAL) v = fp_to_mont(v);
                break;
            case SNARKV_FR_OP_CONST:
                v = fp_load<FR>(consts + (size_t)ins.z * 32);   // converted to Montgomery form by the host entry point
                break;
            case SNARKV_FR_OP_ADD: v = fp_add(fr_load_rw(reg(ins.z)), fr_load_rw(reg(ins.w))); break;
            case SNARKV_FR_OP_SUB: v = fp_sub(fr_load_rw(reg(ins.z)), fr_load_rw(reg(ins.w))); break;
            case SNARKV_FR_OP_MUL: v = fp_mul(fr_load_rw(reg(ins.z)), fr_load_rw(reg(ins.w))); break;
            case SNARKV_FR_OP_NEG: v = fp_neg(fr_load_rw(reg(ins.z))); break;
            case SNARKV_FR_OP_NZ:
                v = fr_load_rw(reg(ins.z));
                if (fp_is_zero(v)) v = fp_one<FR>();
                break;
            case SNARKV_FR_OP_KEEPZ:
                v = fp_is_zero(fr_load_rw(reg(ins.w))) ? fp_zero<FR>() : fr_load_rw(reg(ins.z));
                break;
            default: {   // SNARKV_FR_OP_INV (the host validated the opcodes)
                v = fr_load_rw(reg(ins.z));
                if (!fp_is_zero(v)) v = fp_inv_serial(v);
                break;
            }
        }
        fp_store<FR>(reg(ins.y), v);
    }
    for (uint32_t k = 0; k < n_out; ++k) {
        Fr v = fr_load_rw(reg(out_regs[k]));
        if (format == SNARKV_CANONICAL) v = fp_from_mont(v);
        fp_store<FR>(outputs + (p * n_out + k) * 32, v);
    }
}

// consts arrive in `format`; the kernel wants Montgomery form
__global__ void k_fr_consts_to_mont(uint8_t* __restrict__ consts, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fp_store<FR>(consts + (size_t)i * 32, fp_to_mont(fp_load<FR>(consts + (size_t)i * 32)));
}

int fr_program_device(snarkv_ctx* ctx, const void* d_prog, size_t n_instr, void* d_consts, size_t n_consts, const void* d_inputs,
                      size_t n_inputs, size_t m, int format, uint32_t n_regs, const void* d_out_regs, size_t n_out, void* d_outputs) {
    uint8_t* regs = (uint8_t*)ctx->wsget(WS_FR_REGS, (size_t)n_regs * m * 32);
    if (!regs) return SNARKV_ERR_CUDA;
    Stage sg(ctx, "fr_program");
    if (format == SNARKV_CANONICAL && n_consts) {
        k_fr_consts_to_mont<<<(unsigned)((n_consts + 127) / 128), 128, 0, ctx->stream>>>((uint8_t*)d_consts, (uint32_t)n_consts);
        SNARKV_LAUNCH_CHECK(ctx, "k_fr_consts_to_mont");
        sg.launched();
    }
    k_fr_program<<<(unsigned)((m + 127) / 128), 128, 0, ctx->stream>>>((const uint4*)d_prog, (uint32_t)n_instr, (const uint8_t*)d_consts,
                                                                     (const uint8_t*)d_inputs, (uint32_t)n_inputs, m, format, regs,
                                                                     (const uint32_t*)d_out_regs, (uint32_t)n_out, (uint8_t*)d_outputs);
    SNARKV_LAUNCH_CHECK(ctx, "k_fr_program");
    sg.launched();
    return SNARKV_OK;
}

}  // namespace snarkv
