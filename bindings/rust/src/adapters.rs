//! UNCOMPILED reference text.  The two adapters a DISTINCT loader type needs, because the reference implements its transcripts and
//! its accumulator encoding for `NativeLoader` only:
//!   * `EvmTranscript<C, NativeLoader, S, B>`   snark-verifier/src/system/halo2/transcript/evm.rs:175-268
//!   * `PoseidonTranscript<C, NativeLoader, ..>` system/halo2/transcript/halo2.rs:176-274
//!   * `AccumulatorEncoding<C, NativeLoader> for LimbsEncoding<LIMBS, BITS>`  pcs/kzg/accumulator.rs:50-82
//! `PlonkVerifier::<KzgAs<Bn256, Gwc19>, LimbsEncoding<4, 68>>::read_proof / verify` are generic over `L: Loader`, but with
//! `L = CudaLoader` they need `T: TranscriptRead<G1Affine, CudaLoader>` and `AE: AccumulatorEncoding<G1Affine, CudaLoader>`,
//! which do not exist upstream.  Both are pure delegation: `CudaScalar(Fr)` / `CudaPoint(G1Affine)` are newtypes of exactly the
//! values `NativeLoader` loads (loader/native.rs:44,75), so nothing is recomputed.
//!
//! (If no distinct loader type is wanted, none of this is needed: see in_crate_hook.rs.)
use crate::cuda_loader::{CudaLoader, CudaPoint, CudaScalar};
use halo2curves::bn256::{Fr, G1Affine};
use snark_verifier::{
    loader::native::NativeLoader,
    pcs::{
        kzg::{KzgAccumulator, LimbsEncoding},
        AccumulatorEncoding,
    },
    util::transcript::{Transcript, TranscriptRead},
    Error,
};

/// Wraps ANY native transcript reader (Keccak `EvmTranscript`, `PoseidonTranscript`, halo2's Blake2b adapter) and presents it as a
/// transcript over `CudaLoader`.  util/transcript.rs:9-56 is the whole trait surface.
pub struct CudaTranscript<T>(pub T);

impl<T: Transcript<G1Affine, NativeLoader>> Transcript<G1Affine, CudaLoader> for CudaTranscript<T> {
    fn loader(&self) -> &CudaLoader {
        &CudaLoader
    }
    fn squeeze_challenge(&mut self) -> CudaScalar {
        CudaScalar(self.0.squeeze_challenge())
    }
    fn common_ec_point(&mut self, ec_point: &CudaPoint) -> Result<(), Error> {
        self.0.common_ec_point(&ec_point.0)
    }
    fn common_scalar(&mut self, scalar: &CudaScalar) -> Result<(), Error> {
        self.0.common_scalar(&scalar.0)
    }
}

impl<T: TranscriptRead<G1Affine, NativeLoader>> TranscriptRead<G1Affine, CudaLoader> for CudaTranscript<T> {
    fn read_scalar(&mut self) -> Result<CudaScalar, Error> {
        self.0.read_scalar().map(CudaScalar)
    }
    fn read_ec_point(&mut self) -> Result<CudaPoint, Error> {
        // the native reader has already validated the point (`from_xy` / `from_bytes`: transcript/evm.rs:247-266, halo2.rs:261-272)
        self.0.read_ec_point().map(CudaPoint)
    }
}

/// pcs/kzg/accumulator.rs:50-82 for `CudaLoader`.  A local type (the orphan rule forbids implementing the upstream trait for the
/// upstream `LimbsEncoding` with an upstream-only parameter list... `CudaLoader` is local, so the impl on `LimbsEncoding` itself is
/// also legal; the local wrapper keeps the companion crate independent of that detail).
#[derive(Clone, Debug)]
pub struct CudaLimbsEncoding<const LIMBS: usize, const BITS: usize>;

impl<const LIMBS: usize, const BITS: usize> AccumulatorEncoding<G1Affine, CudaLoader> for CudaLimbsEncoding<LIMBS, BITS> {
    type Accumulator = KzgAccumulator<G1Affine, CudaLoader>;

    fn from_repr(limbs: &[&CudaScalar]) -> Result<Self::Accumulator, Error> {
        let native: Vec<&Fr> = limbs.iter().map(|l| &l.0).collect();
        let acc = <LimbsEncoding<LIMBS, BITS> as AccumulatorEncoding<G1Affine, NativeLoader>>::from_repr(&native)?;
        Ok(KzgAccumulator::new(CudaPoint(acc.lhs), CudaPoint(acc.rhs)))
    }
}

// Usage (replaces INTEGRATION.md's former "same call with L = CudaLoader"):
//
//   let mut transcript = CudaTranscript(EvmTranscript::<G1Affine, NativeLoader, _, _>::new(proof_bytes.as_slice()));
//   let instances: Vec<Vec<CudaScalar>> = instances.iter().map(|col| col.iter().copied().map(CudaScalar).collect()).collect();
//   let protocol: PlonkProtocol<G1Affine, CudaLoader> = protocol.loaded(&CudaLoader);            // verifier/plonk/protocol.rs:112-139
//   type Verifier = PlonkVerifier<KzgAs<Bn256, Gwc19>, CudaLimbsEncoding<4, 68>>;
//   let proof = Verifier::read_proof(&dk, &protocol, &instances, &mut transcript)?;               // verifier/plonk.rs:113-123
//   Verifier::verify(&dk, &protocol, &instances, &proof)?;                                        // MSMs and decide on the GPU
