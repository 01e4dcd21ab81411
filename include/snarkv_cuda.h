/* snarkv_cuda.h — C ABI of libsnarkv_cuda.so: the B200-native replacement for snark-verifier's native
 * verification hot path (BN254 G1 multi-scalar multiplication + KZG accumulator pairing check).
 *
 * Every entry point names the reference interface it replaces (paths relative to
 * /root/reference/snark-verifier/src).  INTEGRATION.md shows the Rust `extern "C"` block and the
 * `impl EcPointLoader / AccumulationDecider for CudaLoader` a maintainer would add on the reference side.
 *
 * There is NO CPU fallback anywhere behind this header: every compute entry point launches sm_100a kernels and
 * returns SNARKV_ERR_CUDA if no usable device / kernel image is present.
 *
 * ---- Byte formats -----------------------------------------------------------------------------------------------
 *  scalar (Fr)      32 B   SNARKV_CANONICAL : little-endian canonical integer  == `PrimeField::to_repr()`
 *                                              (util/msm.rs:264, util/arithmetic.rs:260)
 *                          SNARKV_MONTGOMERY: halo2curves' in-memory `Fr([u64;4])` limbs (value * 2^256 mod r)
 *                                              == `SerdeObject::to_raw_bytes()`
 *  G1 affine        64 B   x || y, each coordinate in the selected format; the identity is (0,0) (halo2curves)
 *  G1 Jacobian      96 B   X || Y || Z always MONTGOMERY limbs (halo2curves `G1` layout); identity Z == 0
 *  G2 affine       128 B   x.c0 || x.c1 || y.c0 || y.c1, CANONICAL; identity all-zero
 *  GT (Fq12)       384 B   12 x Fq CANONICAL in tower order c0.c0.c0, c0.c0.c1, c0.c1.c0, c0.c1.c1, c0.c2.c0, c0.c2.c1,
 *                          c1.c0.c0 ... c1.c2.c1  with Fq2 = Fq[i]/(i^2+1), Fq6 = Fq2[v]/(v^3-(9+i)), Fq12 = Fq6[w]/(w^2-v)
 *
 * ---- Errors -------------------------------------------------------------------------------------------------------
 *  Return value 0 = ok, < 0 = error (see enum).  A pairing check that REJECTS is data (`accept[i] = 0`), not an error;
 *  the Rust glue maps it to `Error::AssertionFailure("e(lhs, g2)·e(rhs, -s_g2) == O")` (pcs/kzg/decider.rs:81).
 *  n == 0 for an MSM is SNARKV_ERR_EMPTY, mirroring the `.unwrap()` panic at loader/native.rs:69.
 *
 * ---- Threading ------------------------------------------------------------------------------------------------------
 *  A context is thread-compatible, not thread-safe: one call at a time per context (it owns one stream and one grow-only
 *  workspace).  Create one context per host thread / per GPU.  `NativeLoader` is a ZST reachable from any thread
 *  (loader/native.rs:11-19); the Rust glue keeps one lazily created context per thread to offer the same.
 */
#ifndef SNARKV_CUDA_H
#define SNARKV_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct snarkv_ctx snarkv_ctx;

enum {
    SNARKV_OK = 0,
    SNARKV_ERR_USAGE = -1,       /* null pointer, bad enum, size overflow */
    SNARKV_ERR_EMPTY = -2,       /* empty MSM (native.rs:69 panics) */
    SNARKV_ERR_CUDA = -3,        /* CUDA runtime / no device / no kernel image; see snarkv_last_error */
    SNARKV_ERR_BAD_SCALAR = -4,  /* CANONICAL scalar >= r (`from_repr` would fail) */
    SNARKV_ERR_BAD_POINT = -5,   /* coordinate >= p, or point not on the curve (`from_xy` would fail, accumulator.rs:75-78) */
    SNARKV_ERR_NO_KEY = -6       /* decide called before snarkv_kzg_set_deciding_key */
};

enum { SNARKV_CANONICAL = 0, SNARKV_MONTGOMERY = 1 };

/* Validation of MSM inputs.  halo2curves values are valid by construction, so the Rust glue passes 0;
 * untrusted byte sources pass SNARKV_CHECK_INPUTS to get from_repr / from_xy behaviour on the device. */
enum { SNARKV_CHECK_INPUTS = 1 };

/* ---- context ---------------------------------------------------------------------------------------------------------
 * Replaces the global `LOADER: NativeLoader` (loader/native.rs:11-15).  `device` is a CUDA ordinal. */
int snarkv_init(int device, snarkv_ctx** out);
void snarkv_destroy(snarkv_ctx* ctx);
const char* snarkv_last_error(const snarkv_ctx* ctx); /* never NULL; "" when no error */
const char* snarkv_version(void);
/* Launch all subsequent work of this context on an external stream (a `cudaStream_t`, e.g. torch's current stream);
 * NULL restores the context's own stream. */
int snarkv_set_stream(snarkv_ctx* ctx, void* cuda_stream);
/* MSM tuning: window bits c (0 = choose from n). */
int snarkv_set_window_bits(snarkv_ctx* ctx, int c);

/* MSM tuning: GLV endomorphism split of every term into two half-length terms (halves the windows): 0 = for n < 2^22 (default),
 * 1 = always, 2 = never.  Results are identical. */
int snarkv_set_glv_mode(snarkv_ctx* ctx, int mode);

/* MSM tuning: bucket accumulation kernel (util/msm.rs:291-296).  0 = choose from the mean bucket load (default), 1 = XYZZ mixed
 * additions (10 multiplications per point), 2 = batched affine additions with a shared inversion per <= 4096 additions (6 per
 * point), 3 = run both and compare every task result on the device (debugging aid; synchronous).  Results are identical. */
int snarkv_set_accumulate_mode(snarkv_ctx* ctx, int mode);

/* KZG decide tuning: 0 = choose from N (default), 1 = one thread per check (largest batches), 2 = cooperative (block or warp
 * per check, chosen from N), 3 = one 160-thread block per check (lowest latency), 4 = one warp per check.  Results are identical. */
int snarkv_set_pairing_mode(snarkv_ctx* ctx, int mode);

/* ---- a1: EcPointLoader::multi_scalar_multiplication --------------------------------------------------------------------
 * Replaces `NativeLoader::multi_scalar_multiplication(pairs) -> C` (loader.rs:108-113, loader/native.rs:61-71):
 *   out = to_affine( sum_i points[i] * scalars[i] ).
 * Host buffers in (n x 32 B scalars, n x 64 B affine points, both in `format`), 64-byte affine result out (same format).
 * The algorithm behind it is a signed-digit Pippenger (the reference's own large-n algorithm is util/msm.rs:259-343);
 * the affine result is canonical, hence bit-identical to the reference fold. */
int snarkv_g1_msm(snarkv_ctx* ctx, const uint8_t* scalars, const uint8_t* points, size_t n, int format, int flags,
                  uint8_t out_affine[64]);

/* The same MSM against a RESIDENT base set.  A verifier's fixed bases (the preprocessed / vk commitments of a `PlonkProtocol`,
 * verifier/plonk/protocol.rs:30-34, the SRS generator of `KzgSuccinctVerifyingKey`, pcs/kzg.rs:19-35) do not change between calls:
 * upload them once — validated (SNARKV_CHECK_INPUTS), converted and, for the GLV plans, extended by their endomorphism images on the
 * device — and pass only the n x 32 B scalars per call (one scalar per base, in the order of the upload).  Results are identical to
 * snarkv_g1_msm on the same pairs.  The handle belongs to the context's device and must be freed before the context. */
typedef struct snarkv_bases snarkv_bases;
int snarkv_g1_bases_upload(snarkv_ctx* ctx, const uint8_t* points, size_t n, int format, int flags, snarkv_bases** out);
void snarkv_g1_bases_free(snarkv_ctx* ctx, snarkv_bases* bases);
int snarkv_g1_msm_bases_resident(snarkv_ctx* ctx, const snarkv_bases* bases, const uint8_t* scalars, size_t n, int format, int flags,
                                 uint8_t out_affine[64]);

/* One rank's share of a chunk-partitioned MSM (util/msm.rs:322-332: `scalars.chunks(chunk_size).zip(bases.chunks(..))`, one
 * serial MSM per chunk): host slices in, the chunk's Jacobian partial (96 B) left in DEVICE memory, ready for the
 * all-gather + snarkv_g1_fold_partials_device that replaces the fold at util/msm.rs:333-335.  Synchronises the stream. */
int snarkv_g1_msm_partial(snarkv_ctx* ctx, const uint8_t* scalars, const uint8_t* points, size_t n, int format, int flags,
                          void* d_out_jacobian);

/* Same operation on operands already resident in this device's HBM (pointers from cudaMalloc / a torch tensor's
 * data_ptr()).  Results are written to DEVICE memory, nothing is synchronised or copied to the host:
 *   d_out_affine   64 B  affine result in `format`            (may be NULL)
 *   d_out_jacobian 96 B  the same point as a Jacobian partial   (may be NULL) — the unit exchanged between GPUs.
 * `d_status` (4 B, device, may be NULL) receives 0 or the first SNARKV_ERR_BAD_* found when SNARKV_CHECK_INPUTS is set. */
int snarkv_g1_msm_device(snarkv_ctx* ctx, const void* d_scalars, const void* d_points, size_t n, int format, int flags,
                         void* d_out_affine, void* d_out_jacobian, void* d_status);

/* Fold k Jacobian partials (k x 96 B, device memory) into one affine point (64 B, device memory, `format`):
 * the `results.iter().fold(identity, |acc, r| acc + r)` + `.to_affine()` of util/msm.rs:333-335 across GPUs,
 * applied after an all-gather of the per-rank partials. */
int snarkv_g1_fold_partials_device(snarkv_ctx* ctx, const void* d_partials, size_t k, int format, void* d_out_affine);

/* m independent MSMs in one launch: MSM j covers terms [offsets[j], offsets[j+1]) of the scalar/point arrays
 * (offsets has m+1 entries, offsets[0] == 0).  out = m x 64 B.  This is PlonkVerifier running `Msm::evaluate`
 * (util/msm.rs:81-98) for many proofs at once; every segment must be non-empty. */
int snarkv_g1_msm_batch(snarkv_ctx* ctx, const uint8_t* scalars, const uint8_t* points, const uint64_t* offsets, size_t m,
                        int format, int flags, uint8_t* out_affine);

/* The same m MSMs fused into ONE by random linear combination:  out = sum_j rho^j * MSM_j  — the batching of
 * pcs/kzg/decider.rs:146-185 applied to the verifier's `Msm::evaluate` calls themselves, so that a batch of proofs costs one
 * large Pippenger pass instead of m small ones.  Scalars are multiplied by rho^j on the device (`r.powers(m)`, loader.rs:71-78). */
int snarkv_g1_msm_batch_rlc(snarkv_ctx* ctx, const uint8_t* scalars, const uint8_t* points, const uint64_t* offsets, size_t m,
                            const uint8_t rho[32], int format, int flags, uint8_t out_affine[64]);

/* ---- a6: KzgAs::verify (accumulate) ------------------------------------------------------------------------------------
 * Replaces pcs/kzg/accumulation.rs:41-63: powers_of_r = r.powers(n) (loader.rs:71-78);
 *   out_lhs = sum_i r^i lhs[i], out_rhs = sum_i r^i rhs[i]   (two MSMs; a blinding pair, if any, is just the last pair).
 * lhs/rhs: n x 64 B, r: 32 B, outputs 64 B each, all in `format`.  The powers are produced on the device. */
int snarkv_kzg_accumulate(snarkv_ctx* ctx, const uint8_t* lhs, const uint8_t* rhs, size_t n, const uint8_t r[32], int format,
                          uint8_t out_lhs[64], uint8_t out_rhs[64]);

/* ---- a7/a8/a9: KzgDecidingKey + AccumulationDecider::{decide, decide_all} -----------------------------------------------
 * `KzgDecidingKey::new(svk.g, g2, s_g2)` (pcs/kzg/decider.rs:6-42).  Precomputes `G2Prepared::from(g2)` and
 * `G2Prepared::from(-s_g2)` once (the reference recomputes both inside every decide call, decider.rs:74).
 * g1: 64 B CANONICAL, g2 / s_g2: 128 B CANONICAL. */
int snarkv_kzg_set_deciding_key(snarkv_ctx* ctx, const uint8_t g1[64], const uint8_t g2[128], const uint8_t s_g2[128]);

/* `KzgAs::decide` for N accumulators at once (decider.rs:70-93):
 *   accept[i] = ( e(lhs_i, g2) * e(rhs_i, -s_g2) == 1 ).
 * lhs/rhs: N x 64 B in `format`; accept: N bytes; gt_out: N x 384 B CANONICAL or NULL.
 * `decide` is N = 1; `decide_all` is "all accept[i] == 1" (the reference stops at the first failure — callers that
 * need that index take the first zero). */
int snarkv_kzg_decide_batch(snarkv_ctx* ctx, const uint8_t* lhs, const uint8_t* rhs, size_t N, int format, uint8_t* accept,
                            uint8_t* gt_out);
/* Device-resident variant: d_lhs/d_rhs N x 64 B, d_accept N bytes, d_gt N x 384 B or NULL; no host synchronisation. */
int snarkv_kzg_decide_batch_device(snarkv_ctx* ctx, const void* d_lhs, const void* d_rhs, size_t N, int format, void* d_accept,
                                   void* d_gt);

/* ---- a11: batched decision by random linear combination -------------------------------------------------------------------
 * The reference batches pairing checks this way in the EVM loader's `decide_all` (pcs/kzg/decider.rs:146-185): with powers of a
 * challenge rho,  lhs' = sum_i rho^i lhs_i,  rhs' = sum_i rho^i rhs_i  (two MSMs), then ONE `decide(lhs', rhs')`.
 * All N accumulators valid  =>  *accept = 1;  any invalid one  =>  *accept = 0 except with probability ~N/r over rho.
 * rho must be unpredictable to whoever produced the accumulators (the reference hashes all the points, decider.rs:164-168;
 * hashing stays on the host).  lhs/rhs N x 64 B and rho 32 B in `format`; out_lhs/out_rhs (64 B each, may be NULL) return the
 * combined accumulator.  Unlike snarkv_kzg_decide_batch this cannot say WHICH accumulator failed. */
int snarkv_kzg_decide_all_fused(snarkv_ctx* ctx, const uint8_t* lhs, const uint8_t* rhs, size_t N, const uint8_t rho[32], int format,
                                uint8_t* accept, uint8_t out_lhs[64], uint8_t out_rhs[64]);

/* ---- (next row f1) Fr scalar preparation — the step right before the MSM -----------------------------------------------------
 * `LoadedScalar::powers(n)` (loader.rs:71-78): out[i] = r^i, i < n.   n x 32 B, `format` in and out. */
int snarkv_fr_powers(snarkv_ctx* ctx, const uint8_t r[32], size_t n, int format, uint8_t* out);
/* `ScalarLoader::batch_invert` (loader.rs:255-262) / `batch_invert_and_mul` (util/arithmetic.rs:47-69), in place: every non-zero
 * value v becomes coeff / v (coeff = 1 when NULL); zeros are left untouched, exactly as the reference does. */
int snarkv_fr_batch_invert(snarkv_ctx* ctx, uint8_t* values, size_t n, const uint8_t* coeff, int format);
/* Element-wise product out[i] = a[i] * b[i] (the RLC scalars rho^i * s_i of the fused batch paths). */
int snarkv_fr_mul_vec(snarkv_ctx* ctx, const uint8_t* a, const uint8_t* b, size_t n, int format, uint8_t* out);

/* ---- (next row f2) Fiat-Shamir challenges of the Keccak `EvmTranscript` for m proofs of one transcript shape -------------------
 * Replaces system/halo2/transcript/evm.rs:184-222 + loader/evm/util.rs:61-67 for a batch.  `streams`: m x stream_len bytes, the
 * 32-byte big-endian words each proof absorbs, in order (common_scalar: 1 word, common_ec_point: x then y).  `seg_end[i]`: byte
 * offset (multiple of 32, non-decreasing) after which the i-th challenge is squeezed.  Semantics per proof:
 *   H_i = Keccak256( H_{i-1} || stream[seg_end[i-1] .. seg_end[i]) [|| 0x01 when that is exactly 32 bytes] ),  challenge_i = be(H_i) mod r.
 * `challenges`: m x k x 32 B scalars in `format`. */
int snarkv_evm_transcript_challenges(snarkv_ctx* ctx, const uint8_t* streams, size_t stream_len, const uint32_t* seg_end, size_t k, size_t m,
                                     int format, uint8_t* challenges);

/* ---- (next row f2) Poseidon transcript challenges and compressed proof points for m proofs of one transcript shape --------------
 * Replaces util/hash/poseidon.rs:117-203 (Poseidon<F, L, 5, 4> with R_F = 8, R_P = 60: the instance snark-verifier-sdk/src/halo2.rs:53-56
 * uses) and the native PoseidonTranscript of system/halo2/transcript/halo2.rs:201-274 for a batch.
 *   snarkv_poseidon_transcript_challenges  `elements`: m x stream_len scalar-field elements (32 B each, `format`) in the order each
 *       proof absorbs them — common_scalar = 1 element, common_ec_point = (x mod r, y mod r); `seg_end[i]`: ELEMENT offset after which
 *       the i-th challenge is squeezed.  Per squeeze the buffered elements are absorbed RATE at a time (a padding 1 behind the last
 *       one, one extra permutation when their number is a multiple of RATE) and state[1] is the challenge.  `challenges`: m x k x 32 B.
 *   snarkv_g1_decompress_batch  `C::from_bytes` (read_ec_point, halo2.rs:258-272) for n compressed bn256 G1 points: 32 B = x little-endian,
 *       bit 255 = parity of y, bit 254 = identity flag.  `points`: n x 64 B affine in `format`; `fr_elements` (may be NULL): n x 2 x 32 B,
 *       (x mod r, y mod r) — what common_ec_point absorbs; `valid[i]` = 0 for a non-canonical x, a point off the curve or the
 *       identity (which the reference's transcript rejects as well, halo2.rs:226-241).
 *   snarkv_poseidon_permute  parity entry: m states of 5 elements -> permuted states (pins the device permutation against the public
 *       Poseidon test vectors). */
int snarkv_poseidon_transcript_challenges(snarkv_ctx* ctx, const uint8_t* elements, size_t stream_len, const uint32_t* seg_end, size_t k, size_t m,
                                          int format, uint8_t* challenges);
int snarkv_g1_decompress_batch(snarkv_ctx* ctx, const uint8_t* compressed, size_t n, int format, uint8_t* points, uint8_t* fr_elements,
                               uint8_t* valid);
int snarkv_poseidon_permute(snarkv_ctx* ctx, const uint8_t* states, size_t m, int format, uint8_t* out);

/* ---- (row a13) `LimbsEncoding<LIMBS, BITS>::from_repr` for m accumulators -----------------------------------------------------
 * Replaces pcs/kzg/accumulator.rs:57-81 (+ util/arithmetic.rs:270-282 fe_from_limbs): `limbs` = m x 4 x num_limbs x 32 B scalars in
 * `format`, per accumulator the limbs of lhs.x, lhs.y, rhs.x, rhs.y (limb i weighs 2^(limb_bits i); the SDK uses 4 x 68).
 * `lhs`, `rhs`: m x 64 B affine points in `format`; `valid[a]` = 1 iff every coordinate fits 32 bytes, is a canonical base-field
 * element and both points satisfy `from_xy` (on the curve, or the identity (0,0)) — where the reference panics, valid[a] = 0 and
 * the points are written as (0,0).  1 <= num_limbs <= 8, limb_bits (num_limbs - 1) <= 256. */
int snarkv_kzg_accumulators_from_limbs(snarkv_ctx* ctx, const uint8_t* limbs, size_t m, uint32_t num_limbs, uint32_t limb_bits, int format,
                                       uint8_t* lhs, uint8_t* rhs, uint8_t* valid);

/* ---- (next row f3) per-proof PLONK scalar evaluation for m proofs of one protocol -----------------------------------------
 * Replaces verifier/plonk/protocol.rs:211-283 (CommonPolynomialEvaluation), :333-392 (Expression::evaluate) and
 * verifier/plonk/proof.rs:298-349 (instance evaluations, quotient evaluation) for a batch: the protocol — hence the expression
 * tree — is the same for every proof, so the host flattens it ONCE into a straight-line register program and one GPU thread per
 * proof runs it.  Instruction (op, dst, a, b), registers hold Fr values:
 *   INPUT dst <- inputs[proof][a]   CONST dst <- consts[a]   ADD/SUB/MUL dst <- reg[a] (op) reg[b]   NEG dst <- -reg[a]
 *   INV   dst <- 1 / reg[a], a zero stays zero (ScalarLoader::batch_invert, loader.rs:255-262 / util/arithmetic.rs:47-74).
 *   NZ    dst <- reg[a] == 0 ? 1 : reg[a]       KEEPZ dst <- reg[b] == 0 ? 0 : reg[a]
 *         (the two selects with which the host emits util/arithmetic.rs:47-69 `batch_invert` — skip the zeros, ONE inversion for
 *         all denominators of a proof — as straight-line code)
 * `inputs`: m x n_inputs x 32 B, `consts`: n_consts x 32 B, `outputs`: m x n_out x 32 B (register out_regs[k] of every proof),
 * all in `format`.  Every register must be written before it is read (checked); dst, a, b < n_regs. */
enum { SNARKV_FR_OP_INPUT = 0, SNARKV_FR_OP_CONST = 1, SNARKV_FR_OP_ADD = 2, SNARKV_FR_OP_SUB = 3, SNARKV_FR_OP_MUL = 4,
       SNARKV_FR_OP_NEG = 5, SNARKV_FR_OP_INV = 6, SNARKV_FR_OP_NZ = 7, SNARKV_FR_OP_KEEPZ = 8 };
typedef struct snarkv_fr_instr { uint32_t op, dst, a, b; } snarkv_fr_instr;
int snarkv_fr_program_eval_batch(snarkv_ctx* ctx, const snarkv_fr_instr* program, size_t n_instr, uint32_t n_regs, const uint8_t* consts,
                                 size_t n_consts, const uint8_t* inputs, size_t n_inputs, size_t m, const uint32_t* out_regs, size_t n_out,
                                 int format, uint8_t* outputs);
/* The same with inputs and outputs resident in device memory (program, consts and out_regs are still host arrays: they are small
 * and validated on the host). */
int snarkv_fr_program_eval_batch_device(snarkv_ctx* ctx, const snarkv_fr_instr* program, size_t n_instr, uint32_t n_regs,
                                        const uint8_t* consts, size_t n_consts, const void* d_inputs, size_t n_inputs, size_t m,
                                        const uint32_t* out_regs, size_t n_out, int format, void* d_outputs);

/* ---- multi-GPU: one host call, all the GPUs of the box -----------------------------------------------------------------------
 * What the rayon `parallel` feature is to the reference's large MSM (util/msm.rs:311-336: one contiguous chunk of the term slice
 * per thread, partial results folded), a multi-device context is here: device g gets terms [g ceil(n/G), (g+1) ceil(n/G)), runs the
 * single-device pipeline on them from its own host thread and leaves a 96-byte Jacobian partial in its HBM; the single exchange
 * step (the fold of util/msm.rs:333-335) is ONE kernel on device 0 that reads every partial out of its producer's memory over
 * NVLink peer access.  This is the entry a `CudaLoader::multi_scalar_multiplication` (loader.rs:108-113) binds to use the whole
 * box.  Units that are independent shard with no exchange: pairing checks (decider.rs:84-93) in contiguous blocks, RLC-fused
 * batches by MSM segment (device g starts its powers of rho at rho^(first segment of g)).  Results are bit-identical to the
 * single-device entry points for every device count.
 * `devices` = CUDA ordinals (NULL: 0 .. n_devices - 1; n_devices <= 0: every visible device).  Peer access from devices[0] to
 * the others is required (SNARKV_ERR_CUDA otherwise).  Thread-compatible like snarkv_ctx; snarkv_multi_ctx(m, i) exposes the
 * per-device context for the tuning setters. */
typedef struct snarkv_multi snarkv_multi;
int snarkv_multi_init(const int* devices, int n_devices, snarkv_multi** out);
void snarkv_multi_destroy(snarkv_multi* m);
int snarkv_multi_device_count(const snarkv_multi* m);
snarkv_ctx* snarkv_multi_ctx(snarkv_multi* m, int i);
const char* snarkv_multi_last_error(const snarkv_multi* m);
int snarkv_multi_g1_msm(snarkv_multi* m, const uint8_t* scalars, const uint8_t* points, size_t n, int format, int flags,
                        uint8_t out_affine[64]);
int snarkv_multi_g1_msm_batch_rlc(snarkv_multi* m, const uint8_t* scalars, const uint8_t* points, const uint64_t* offsets, size_t segs,
                                  const uint8_t rho[32], int format, int flags, uint8_t out_affine[64]);
int snarkv_multi_kzg_set_deciding_key(snarkv_multi* m, const uint8_t g1[64], const uint8_t g2[128], const uint8_t s_g2[128]);
int snarkv_multi_kzg_decide_batch(snarkv_multi* m, const uint8_t* lhs, const uint8_t* rhs, size_t N, int format, uint8_t* accept,
                                  uint8_t* gt_out);

/* ---- BASELINE config 3 in ONE call: m proofs of one PLONK protocol, from their absorbed byte streams to the fused accumulator ------
 * Replaces, for a batch, PlonkProof::read (verifier/plonk/proof.rs:52-169), PlonkSuccinctVerifier::verify (verifier/plonk.rs:57-93: the
 * per-proof Fr work, as the program snark_verifier_b200.plonk.compile_plonk_verifier emits for the protocol), the two `Msm::evaluate`
 * calls of the multi-open verifier (util/msm.rs:81-98) fused by powers of rho (pcs/kzg/decider.rs:146-185) and, with `decide`, the
 * pairing check (decider.rs:70-82).  Everything between the upload of the streams and the 128-byte result stays in HBM: transcript
 * challenges, parsing (32-byte big-endian words -> scalars, rejected when >= r like `read_scalar`; points validated like `from_xy`),
 * the scalar program, one Pippenger pass per side, one pairing.
 * `streams`: m x stream_words x 32 B = what each proof's Keccak EvmTranscript absorbs, in order: [initial state] | instances | proof.
 * Plan description (all arrays are copied): `seg_end[i]` = WORD offset after which challenge i is squeezed; `row_src[i]` >= 0: program
 * input i is stream word row_src[i] (`row_check[i]` = 1: it is a proof scalar and must be canonical), < 0: challenge -(row_src[i] + 1);
 * `lhs_src` / `rhs_src`[k] >= 0: base k of that MSM is the proof point at words src, src + 1 (x, y), < 0: constant point -(src + 1) of
 * `const_points` (canonical little-endian x || y: SRS generator, preprocessed commitments); program outputs = lhs scalars then rhs scalars.
 * transcript = 1 (native PoseidonTranscript, system/halo2/transcript/halo2.rs:201-274): `streams` = m x (n_pre + n_items) x 32 B — per proof n_pre
 * little-endian scalars (initial state, instances) followed by the proof's n_items 32-byte items (little-endian scalar or compressed point).  Item i is a
 * scalar absorbed as element item_off[i] >= 0, or a point (item_off[i] = -(off + 1): its two elements go to off, off + 1; item_pt[i] = its index among
 * the proof's n_points points).  `stream_words` / `seg_end` / `row_src` count ELEMENTS, `lhs_src` / `rhs_src` >= 0 are point indices.
 * Errors: SNARKV_ERR_BAD_SCALAR / SNARKV_ERR_BAD_POINT when any proof of the batch carries an invalid encoding (Error::Transcript). */
typedef struct snarkv_plonk_plan snarkv_plonk_plan;
typedef struct {
    uint32_t transcript;   /* 0 = Keccak EvmTranscript, 1 = Poseidon transcript (T = 5, RATE = 4, R_F = 8, R_P = 60) */
    uint32_t stream_words;
    const uint32_t* seg_end;
    uint32_t n_challenges;
    const snarkv_fr_instr* program;
    size_t n_instr;
    uint32_t n_regs;
    const uint8_t* consts; /* n_consts x 32 B canonical little-endian */
    size_t n_consts;
    uint32_t n_inputs;
    const uint32_t* out_regs;
    uint32_t n_out;
    const int32_t* row_src;
    const uint8_t* row_check;
    uint32_t n_lhs, n_rhs;
    const int32_t* lhs_src;
    const int32_t* rhs_src;
    const uint8_t* const_points;
    uint32_t n_const_points;
    /* transcript 1 only (see below): leading caller-supplied elements, proof items, item table, points per proof */
    uint32_t n_pre, n_items;
    const int32_t* item_off;
    const int32_t* item_pt;
    uint32_t n_points;
} snarkv_plonk_plan_desc;
int snarkv_plonk_plan_create(snarkv_ctx* ctx, const snarkv_plonk_plan_desc* desc, snarkv_plonk_plan** out);
void snarkv_plonk_plan_free(snarkv_ctx* ctx, snarkv_plonk_plan* plan);
/* -> out_lhs / out_rhs (64 B canonical, may be NULL) = sum_j rho^j (accumulator of proof j); decide != 0: accept[0] = 1 iff it decides */
int snarkv_plonk_accumulate_batch(snarkv_ctx* ctx, snarkv_plonk_plan* plan, const uint8_t* streams, size_t m, const uint8_t rho[32], int decide,
                                  uint8_t out_lhs[64], uint8_t out_rhs[64], uint8_t* accept);

/* ---- (next row f4) Pallas: the IPA decider and the large MSM behind it ---------------------------------------------------
 * The reference's only in-tree consumer of a 2^k-term MSM is `IpaAs::decide` over the Pasta curves (pcs/ipa/decider.rs:47-70):
 *     h = h_coeffs(&xi, 1)  (pcs/ipa.rs:401-417);   accept  <=>  u == util::msm::multi_scalar_multiplication(&h, &dk.g).to_affine()
 * (util/msm.rs:259-343, the Pippenger this library restates for BN254 as well).  The same pipeline is compiled a second time over
 * Pallas (y^2 = x^3 + 5, generator (-1, 2); csrc/msm_pasta.cu).  Byte formats as above with Fq = Pallas base field, Fr = Pallas
 * scalar field: scalars n x 32 B, affine points n x 64 B, identity (0, 0); SNARKV_MONTGOMERY = halo2curves' in-memory limbs.
 *   snarkv_pallas_msm / _device   Sum scalar_i * point_i over Pallas (host / device operands; `_device` like snarkv_g1_msm_device)
 *   snarkv_pallas_h_coeffs        out[j] = scalar * prod_{bit i of j} xi[k - 1 - i], 2^k values (parity entry; decide does this on the device)
 *   snarkv_ipa_set_deciding_key   `IpaDecidingKey::g` (decider.rs:3-16): 2^k points uploaded ONCE, validated with SNARKV_CHECK_INPUTS
 *   snarkv_ipa_decide_batch       N accumulators (u: N x 64 B, xi: N x k x 32 B, SNARKV_CANONICAL) -> accept[a] = 1 iff accumulator a decides;
 *                                 rejection is data — the glue maps it to Error::AssertionFailure("U == commit(G, h)") (decider.rs:57) */
int snarkv_pallas_msm(snarkv_ctx* ctx, const uint8_t* scalars, const uint8_t* points, size_t n, int format, int flags, uint8_t out_affine[64]);
int snarkv_pallas_msm_device(snarkv_ctx* ctx, const void* d_scalars, const void* d_points, size_t n, int format, int flags, void* d_out_affine,
                             void* d_out_jacobian, void* d_status);
int snarkv_pallas_h_coeffs(snarkv_ctx* ctx, const uint8_t* xi, size_t k, const uint8_t scalar[32], int format, uint8_t* out);
int snarkv_ipa_set_deciding_key(snarkv_ctx* ctx, const uint8_t* g, size_t n, int format, int flags);
int snarkv_ipa_decide_batch(snarkv_ctx* ctx, const uint8_t* u, const uint8_t* xi, size_t k, size_t N, int format, uint8_t* accept);
/* synthetic workload and field-op parity entry over the Pallas fields (same definitions as the BN254 ones below, G = (-1, 2)) */
int snarkv_pallas_synth_scalars_device(snarkv_ctx* ctx, uint64_t seed, uint64_t start, size_t n, int format, void* d_out);
int snarkv_pallas_synth_points_device(snarkv_ctx* ctx, uint64_t seed, uint64_t start, size_t n, int format, void* d_out);
int snarkv_pallas_debug_field_op(snarkv_ctx* ctx, int field, int op, const uint8_t* a, const uint8_t* b, size_t n, uint8_t* out);

/* ---- synthetic workload (bench / tests) ----------------------------------------------------------------------------------
 * Deterministic inputs (the test suite restates the same definition independently):
 *   scalar_i: 4 x splitmix64 limbs, top limb masked to 62 bits, one conditional subtraction of r;
 *   point_i = [t_i] G with t_i = splitmix64(...) | 1 (64-bit), G = (1, 2).
 * Written to DEVICE memory in `format` (n x 32 B / n x 64 B), indices start .. start + n - 1. */
int snarkv_synth_scalars_device(snarkv_ctx* ctx, uint64_t seed, uint64_t start, size_t n, int format, void* d_out);
int snarkv_synth_points_device(snarkv_ctx* ctx, uint64_t seed, uint64_t start, size_t n, int format, void* d_out);

/* ---- test support --------------------------------------------------------------------------------------------------------
 * Element-wise field operation on n CANONICAL 32-byte values (host buffers): field 0 = Fq, 1 = Fr;
 * op 0 = a*b, 1 = a+b, 2 = a-b, 3 = a^-1, 4 = a^2.  Exists so the parity suite can pin the device Montgomery arithmetic
 * (fp.cuh) directly against the golden field vectors; replaces nothing in the reference. */
int snarkv_debug_field_op(snarkv_ctx* ctx, int field, int op, const uint8_t* a, const uint8_t* b, size_t n, uint8_t* out);

/* ---- instrumentation -----------------------------------------------------------------------------------------------------
 * When enabled, every pipeline stage of the next MSM / decide call is bracketed by CUDA events on the context's stream.
 * snarkv_profile_read synchronises and returns up to `cap` (name, milliseconds, launches) records of the LAST call. */
/* The MSM plan the library would use for n terms: out = {window bits c, windows W, buckets per window 2^(c-1), max points
 * per accumulate task}.  Pure host-side query (no launch). */
int snarkv_g1_msm_plan(snarkv_ctx* ctx, size_t n, uint32_t out[4]);
typedef struct {
    const char* name; /* static string, e.g. "msm_bucket_accumulate" */
    float ms;
    int launches;
} snarkv_stage_time;
int snarkv_profile_enable(snarkv_ctx* ctx, int on);
int snarkv_profile_read(snarkv_ctx* ctx, snarkv_stage_time* out, int cap);
/* total kernels launched by this context since creation */
uint64_t snarkv_launch_count(const snarkv_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* SNARKV_CUDA_H */
