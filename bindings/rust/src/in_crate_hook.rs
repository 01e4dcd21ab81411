//! UNCOMPILED reference text.  The drop-in that needs NO new loader type: two `cfg(feature = "cuda")` branches inside the reference
//! crate, calling SAFE functions of this companion crate (the `unsafe` FFI stays here; snark-verifier keeps `#![deny(unsafe_code)]`,
//! lib.rs:8).  Every transcript, `LimbsEncoding`, the SDK and user code keep using `NativeLoader` unchanged.
//!
//! Patch to snark-verifier/src/loader/native.rs:61-71:
//!
//!     fn multi_scalar_multiplication(pairs: &[(&C::Scalar, &C)]) -> C {
//!         #[cfg(feature = "cuda")]
//!         if let Some(out) = snark_verifier_cuda::hook::try_msm::<C>(pairs) {
//!             return out;
//!         }
//!         pairs.iter().cloned().map(|(scalar, base)| *base * scalar).reduce(|acc, value| acc + value).unwrap().to_affine()
//!     }
//!
//! Patch to snark-verifier/src/pcs/kzg/decider.rs:84-93 (`decide` at :70-82 calls `decide_all(dk, vec![acc])` under the feature):
//!
//!     fn decide_all(dk: &Self::DecidingKey, accumulators: Vec<KzgAccumulator<M::G1Affine, NativeLoader>>) -> Result<(), Error> {
//!         #[cfg(feature = "cuda")]
//!         if let Some(verdict) = snark_verifier_cuda::hook::try_decide_all::<M>(dk, &accumulators) {
//!             return verdict;
//!         }
//!         accumulators.into_iter().map(|accumulator| Self::decide(dk, accumulator)).try_collect::<_, Vec<_>, _>()?;
//!         Ok(())
//!     }
//!
//! The reference functions are generic (`C: CurveAffine`, `M: MultiMillerLoop`); the hooks return `None` unless the type IS BN254
//! (checked with `TypeId`) — Pasta / other curves fall through to the stock code.
use crate::ffi::*;
use halo2curves::{
    bn256::{Bn256, Fr, G1Affine},
    group::ff::PrimeField,
    pairing::MultiMillerLoop,
    CurveAffine,
};
use snark_verifier::{
    pcs::kzg::{KzgAccumulator, KzgDecidingKey},
    loader::native::NativeLoader,
    Error,
};
use std::any::{Any, TypeId};

/// Below this many terms the PCIe round trip costs more than the CPU fold (a 21-term `Msm::evaluate` takes ~1 ms on one core).
pub const MIN_GPU_TERMS: usize = 64;

pub fn try_msm<C: CurveAffine + 'static>(pairs: &[(&C::Scalar, &C)]) -> Option<C> {
    if TypeId::of::<C>() != TypeId::of::<G1Affine>() || pairs.len() < MIN_GPU_TERMS {
        return None;
    }
    // safe downcasts instead of transmutes: &C -> &dyn Any -> &G1Affine
    let scalars: Vec<Fr> = pairs.iter().map(|(s, _)| *(*s as &dyn Any).downcast_ref::<Fr>().unwrap()).collect();
    let points: Vec<G1Affine> = pairs.iter().map(|(_, p)| *(*p as &dyn Any).downcast_ref::<G1Affine>().unwrap()).collect();
    let out: G1Affine = crate::cuda_loader::msm_raw(&scalars, &points); // snarkv_g1_msm / snarkv_multi_g1_msm, SNARKV_MONTGOMERY
    Some(*(&out as &dyn Any).downcast_ref::<C>().unwrap())
}

pub fn try_decide_all<M: MultiMillerLoop + 'static>(
    dk: &KzgDecidingKey<M>,
    accumulators: &[KzgAccumulator<M::G1Affine, NativeLoader>],
) -> Option<Result<(), Error>> {
    if TypeId::of::<M>() != TypeId::of::<Bn256>() {
        return None;
    }
    let dk = (dk as &dyn Any).downcast_ref::<KzgDecidingKey<Bn256>>().unwrap();
    let accs = (accumulators as &dyn Any).downcast_ref::<&[KzgAccumulator<G1Affine, NativeLoader>]>()?;
    Some(crate::cuda_decider::decide_all_native(dk, accs)) // same body as decide_all over CudaLoader, on raw G1Affine
}
