//! UNCOMPILED reference text.  Mirrors snark-verifier/src/pcs/kzg/decider.rs:62-94 for `CudaLoader`, plus the batch forms.
use crate::{cuda_loader::*, ffi::*};
use halo2curves::bn256::{Bn256, Fr, G1Affine};
use snark_verifier::{
    pcs::{
        kzg::{KzgAccumulator, KzgAs, KzgDecidingKey, LimbsEncoding},
        AccumulationDecider,
    },
    Error,
};
use std::fmt::Debug;

const ASSERTION: &str = "e(lhs, g2)·e(rhs, -s_g2) == O"; // decider.rs:81

impl<MOS: Clone + Debug> AccumulationDecider<G1Affine, CudaLoader> for KzgAs<Bn256, MOS> {
    type DecidingKey = KzgDecidingKey<Bn256>;

    fn decide(dk: &Self::DecidingKey, acc: KzgAccumulator<G1Affine, CudaLoader>) -> Result<(), Error> {
        Self::decide_all(dk, vec![acc]) // decider.rs:70-82
    }

    fn decide_all(dk: &Self::DecidingKey, accs: Vec<KzgAccumulator<G1Affine, CudaLoader>>) -> Result<(), Error> {
        if accs.is_empty() {
            return Ok(());
        }
        let lhs: Vec<G1Affine> = accs.iter().map(|a| a.lhs.0).collect();
        let rhs: Vec<G1Affine> = accs.iter().map(|a| a.rhs.0).collect();
        let mut accept = vec![0u8; accs.len()];
        CTX.with(|c| unsafe {
            ensure_key(c, dk); // snarkv_kzg_set_deciding_key(svk.g, g2, s_g2): both G2Prepared once per key, not per call (decider.rs:74)
            let rc = snarkv_kzg_decide_batch(c.raw, lhs.as_ptr() as _, rhs.as_ptr() as _, accs.len(), SNARKV_MONTGOMERY,
                                             accept.as_mut_ptr(), std::ptr::null_mut());
            assert_eq!(rc, SNARKV_OK);
        });
        // first error aborts in the reference (`try_collect`, decider.rs:84-93); any rejection maps to the same Error
        accept.iter().all(|&a| a == 1).then_some(()).ok_or_else(|| Error::AssertionFailure(ASSERTION.to_string()))
    }
}

/// decider.rs:146-185 (the EVM loader's RLC `decide_all`) natively: accumulate with powers of `rho`, ONE pairing.
pub fn decide_all_fused(dk: &KzgDecidingKey<Bn256>, accs: &[KzgAccumulator<G1Affine, CudaLoader>], rho: Fr) -> Result<(), Error> {
    let lhs: Vec<G1Affine> = accs.iter().map(|a| a.lhs.0).collect();
    let rhs: Vec<G1Affine> = accs.iter().map(|a| a.rhs.0).collect();
    let (mut ok, mut ol, mut or) = (0u8, G1Affine::default(), G1Affine::default());
    CTX.with(|c| unsafe {
        ensure_key(c, dk);
        assert_eq!(SNARKV_OK, snarkv_kzg_decide_all_fused(c.raw, lhs.as_ptr() as _, rhs.as_ptr() as _, accs.len(), &rho as *const Fr as _,
                                                          SNARKV_MONTGOMERY, &mut ok, &mut ol as *mut _ as _, &mut or as *mut _ as _));
    });
    (ok == 1).then_some(()).ok_or_else(|| Error::AssertionFailure(ASSERTION.to_string()))
}

/// pcs/kzg/accumulator.rs:57-81 for a batch: `limbs` = the instance values selected by `protocol.accumulator_indices`, m x 4 x LIMBS.
pub fn from_repr_batch<const LIMBS: usize, const BITS: usize>(limbs: &[Fr]) -> Result<Vec<KzgAccumulator<G1Affine, CudaLoader>>, Error> {
    let _ = LimbsEncoding::<LIMBS, BITS>;
    let m = limbs.len() / (4 * LIMBS);
    let (mut lhs, mut rhs, mut valid) = (vec![G1Affine::default(); m], vec![G1Affine::default(); m], vec![0u8; m]);
    CTX.with(|c| unsafe {
        assert_eq!(SNARKV_OK, snarkv_kzg_accumulators_from_limbs(c.raw, limbs.as_ptr() as _, m, LIMBS as u32, BITS as u32, SNARKV_MONTGOMERY,
                                                                 lhs.as_mut_ptr() as _, rhs.as_mut_ptr() as _, valid.as_mut_ptr()));
    });
    assert!(valid.iter().all(|&v| v == 1)); // the reference's `.unwrap()`s on from_repr / from_xy
    Ok(lhs.into_iter().zip(rhs).map(|(l, r)| KzgAccumulator::new(CudaPoint(l), CudaPoint(r))).collect())
}

unsafe fn ensure_key(c: &Ctx, dk: &KzgDecidingKey<Bn256>) {
    // canonical bytes of (svk.g, g2, s_g2); a real implementation caches the last key per context
    let (g1, g2, s_g2) = crate::bytes::key_bytes(dk);
    assert_eq!(SNARKV_OK, snarkv_kzg_set_deciding_key(c.raw, g1.as_ptr(), g2.as_ptr(), s_g2.as_ptr()));
}
