//! UNCOMPILED reference text.  `CudaLoader`: mirrors snark-verifier/src/loader/native.rs:11-93.
//!
//! `NativeLoader` is a ZST with a global `LOADER` (native.rs:11-19) and `multi_scalar_multiplication` is an associated function
//! without `&self` (loader.rs:108-113), so the device context is a lazily created per-thread global — the role `LOADER` plays.
//! The blanket impls `impl<C> LoadedEcPoint<C> for C { type Loader = NativeLoader }` (native.rs:21-41) pin raw `G1Affine` / `Fr`
//! to `NativeLoader`, hence the two newtypes.
use crate::ffi::*;
use halo2curves::bn256::{Fr, G1Affine};
use snark_verifier::{
    loader::{EcPointLoader, LoadedEcPoint, LoadedScalar, Loader, ScalarLoader},
    Error,
};

pub struct Ctx {
    pub raw: *mut snarkv_ctx,
}
impl Ctx {
    pub fn new(device: i32) -> Self {
        let mut raw = std::ptr::null_mut();
        // no CPU fallback: without an sm_100 device this fails loudly
        assert_eq!(unsafe { snarkv_init(device, &mut raw) }, SNARKV_OK, "snarkv_init failed: no sm_100 GPU");
        Ctx { raw }
    }
}
impl Drop for Ctx {
    fn drop(&mut self) {
        unsafe { snarkv_destroy(self.raw) }
    }
}
thread_local! { pub static CTX: Ctx = Ctx::new(0); } // contexts are thread-compatible, one per host thread

#[derive(Clone, Debug)]
pub struct CudaLoader;
#[derive(Clone, Debug, PartialEq)]
pub struct CudaPoint(pub G1Affine);
#[derive(Clone, Debug, PartialEq)]
pub struct CudaScalar(pub Fr); // + FieldOps / Add / Sub / Mul / Neg forwarding to Fr, exactly like `impl FieldOps for F`

impl LoadedEcPoint<G1Affine> for CudaPoint {
    type Loader = CudaLoader;
    fn loader(&self) -> &CudaLoader {
        &CudaLoader
    }
}
impl LoadedScalar<Fr> for CudaScalar {
    type Loader = CudaLoader;
    fn loader(&self) -> &CudaLoader {
        &CudaLoader
    }
}

impl EcPointLoader<G1Affine> for CudaLoader {
    type LoadedEcPoint = CudaPoint;
    fn ec_point_load_const(&self, value: &G1Affine) -> CudaPoint {
        CudaPoint(*value)
    }
    fn ec_point_assert_eq(&self, annotation: &str, lhs: &CudaPoint, rhs: &CudaPoint) -> Result<(), Error> {
        lhs.eq(rhs).then_some(()).ok_or_else(|| Error::AssertionFailure(annotation.to_string())) // native.rs:50-59
    }
    fn multi_scalar_multiplication(pairs: &[(&CudaScalar, &CudaPoint)]) -> CudaPoint {
        // native.rs:61-71.  `pairs` is a slice of reference pairs: gather into two contiguous arrays (64 + 32 B per term, the bytes
        // the reference's own fold touches).  Fr([u64; 4]) / G1Affine { x: Fq, y: Fq } are handed over as they lie in memory:
        // that is SNARKV_MONTGOMERY, no conversion on either side; the identity is (0, 0) on both sides.
        let scalars: Vec<Fr> = pairs.iter().map(|(s, _)| s.0).collect();
        let points: Vec<G1Affine> = pairs.iter().map(|(_, p)| p.0).collect();
        let mut out = G1Affine::default();
        let rc = CTX.with(|c| unsafe {
            snarkv_g1_msm(c.raw, scalars.as_ptr() as *const u8, points.as_ptr() as *const u8, pairs.len(), SNARKV_MONTGOMERY, 0,
                          &mut out as *mut G1Affine as *mut u8)
        });
        assert_eq!(rc, SNARKV_OK, "snarkv_g1_msm failed"); // rc = SNARKV_ERR_EMPTY is the reference's `.unwrap()` panic, native.rs:69
        CudaPoint(out)
    }
}
impl ScalarLoader<Fr> for CudaLoader {
    type LoadedScalar = CudaScalar;
    fn load_const(&self, value: &Fr) -> CudaScalar {
        CudaScalar(*value)
    }
    fn assert_eq(&self, annotation: &str, lhs: &CudaScalar, rhs: &CudaScalar) -> Result<(), Error> {
        lhs.eq(rhs).then_some(()).ok_or_else(|| Error::AssertionFailure(annotation.to_string())) // native.rs:82-91
    }
}
impl Loader<G1Affine> for CudaLoader {}
