//! UNCOMPILED reference text.  `extern "C"` surface of include/snarkv_cuda.h used by the glue.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct snarkv_ctx {
    _private: [u8; 0],
}
#[repr(C)]
pub struct FrInstr {
    pub op: u32,
    pub dst: u32,
    pub a: u32,
    pub b: u32,
}

pub const SNARKV_OK: c_int = 0;
pub const SNARKV_CANONICAL: c_int = 0; // `to_repr()` little-endian bytes
pub const SNARKV_MONTGOMERY: c_int = 1; // halo2curves in-memory limbs: zero-copy from &[Fr] / &[G1Affine]

#[link(name = "snarkv_cuda")]
extern "C" {
    pub fn snarkv_init(device: c_int, out: *mut *mut snarkv_ctx) -> c_int;
    pub fn snarkv_destroy(ctx: *mut snarkv_ctx);
    pub fn snarkv_last_error(ctx: *const snarkv_ctx) -> *const c_char;

    // loader/native.rs:61-71
    pub fn snarkv_g1_msm(ctx: *mut snarkv_ctx, scalars: *const u8, points: *const u8, n: usize, format: c_int, flags: c_int,
                         out_affine: *mut u8) -> c_int;
    // one rank's share of a chunk-partitioned MSM (util/msm.rs:322-336): Jacobian partial stays on the device
    pub fn snarkv_g1_msm_partial(ctx: *mut snarkv_ctx, scalars: *const u8, points: *const u8, n: usize, format: c_int, flags: c_int,
                                 d_out_jacobian: *mut c_void) -> c_int;
    pub fn snarkv_g1_fold_partials_device(ctx: *mut snarkv_ctx, d_partials: *const c_void, k: usize, format: c_int,
                                          d_out_affine: *mut c_void) -> c_int;
    // m x Msm::evaluate, independent results / fused by powers of rho (pcs/kzg/decider.rs:146-185 applied before the MSM)
    pub fn snarkv_g1_msm_batch(ctx: *mut snarkv_ctx, scalars: *const u8, points: *const u8, offsets: *const u64, m: usize,
                               format: c_int, flags: c_int, out_affine: *mut u8) -> c_int;
    pub fn snarkv_g1_msm_batch_rlc(ctx: *mut snarkv_ctx, scalars: *const u8, points: *const u8, offsets: *const u64, m: usize,
                                   rho: *const u8, format: c_int, flags: c_int, out_affine: *mut u8) -> c_int;

    // pcs/kzg/accumulation.rs:41-63, pcs/kzg/decider.rs:6-42, 70-93, 146-185
    pub fn snarkv_kzg_accumulate(ctx: *mut snarkv_ctx, lhs: *const u8, rhs: *const u8, n: usize, r: *const u8, format: c_int,
                                 out_lhs: *mut u8, out_rhs: *mut u8) -> c_int;
    pub fn snarkv_kzg_set_deciding_key(ctx: *mut snarkv_ctx, g1: *const u8, g2: *const u8, s_g2: *const u8) -> c_int;
    pub fn snarkv_kzg_decide_batch(ctx: *mut snarkv_ctx, lhs: *const u8, rhs: *const u8, n: usize, format: c_int,
                                   accept: *mut u8, gt_out: *mut u8) -> c_int;
    pub fn snarkv_kzg_decide_all_fused(ctx: *mut snarkv_ctx, lhs: *const u8, rhs: *const u8, n: usize, rho: *const u8,
                                       format: c_int, accept: *mut u8, out_lhs: *mut u8, out_rhs: *mut u8) -> c_int;
    // pcs/kzg/accumulator.rs:57-81
    pub fn snarkv_kzg_accumulators_from_limbs(ctx: *mut snarkv_ctx, limbs: *const u8, m: usize, num_limbs: u32, limb_bits: u32,
                                              format: c_int, lhs: *mut u8, rhs: *mut u8, valid: *mut u8) -> c_int;

    // loader.rs:71-78, 255-262; util/arithmetic.rs:47-69
    pub fn snarkv_fr_powers(ctx: *mut snarkv_ctx, r: *const u8, n: usize, format: c_int, out: *mut u8) -> c_int;
    pub fn snarkv_fr_batch_invert(ctx: *mut snarkv_ctx, values: *mut u8, n: usize, coeff: *const u8, format: c_int) -> c_int;
    // system/halo2/transcript/evm.rs:184-222 for m proofs of one transcript shape
    pub fn snarkv_evm_transcript_challenges(ctx: *mut snarkv_ctx, streams: *const u8, stream_len: usize, seg_end: *const u32,
                                            k: usize, m: usize, format: c_int, challenges: *mut u8) -> c_int;
    // verifier/plonk/protocol.rs:211-283, 336-392; verifier/plonk/proof.rs:298-349 for m proofs of one protocol
    pub fn snarkv_fr_program_eval_batch(ctx: *mut snarkv_ctx, program: *const FrInstr, n_instr: usize, n_regs: u32,
                                        consts: *const u8, n_consts: usize, inputs: *const u8, n_inputs: usize, m: usize,
                                        out_regs: *const u32, n_out: usize, format: c_int, outputs: *mut u8) -> c_int;
}
