//! `src/provider/gpu.rs` for a patched nova-snark 0.23.0 (vendored through `[patch.crates-io]`, the mechanism vimz already
//! uses at vimz/Cargo.toml:53-54).  SOURCE ONLY: the image this repository is built in has no cargo / rustc, so this file has
//! never been compiled; it states exactly which calls of the crate are redirected and with which arguments, on top of the
//! `vimz-gpu-sys` crate next to it.  Everything else in nova-snark -- the augmented circuit, bellperson synthesis, the Poseidon
//! RO, RecursiveSNARK's bookkeeping, CompressedSNARK -- stays as published.
//!
//! Redirected (crate-relative paths of nova-snark 0.23.0):
//!   src/provider/pasta.rs, bn256_grumpkin.rs   Group::vartime_multiscalar_mul          -> msm()            (vimz_msm)
//!   src/r1cs.rs   R1CSShape::multiply_vec                                               -> multiply_vec()   (vimz_multiply_vec)
//!   src/r1cs.rs   R1CSShape::commit_T                                                   -> commit_t()       (vimz_commit_T)
//!   src/r1cs.rs   RelaxedR1CSWitness::fold                                              -> fold_witness()   (vimz_fold_witness)
//!   src/nifs.rs   NIFS::prove (resident fast path, optional)                            -> ResidentFold     (vimz_acc_*)
use std::collections::HashMap;
use std::sync::{Mutex, OnceLock};

use vimz_gpu_sys as sys;
use vimz_gpu_sys::{Affine, Point, Scalar};

/// One GPU context per curve of the cycle, created on first use (device from VIMZ_GPU_DEVICE, default 0).
pub struct CurveGpu {
    pub ctx: sys::Context,
    /// Commitment keys by (address, length) of the `Vec<Affine>` inside `CommitmentKey<G>`: PublicParams owns the vector for the
    /// whole proof, so every `commit` after the first finds its resident window table.
    keys: HashMap<(usize, usize), sys::CommitmentKey>,
}

static CURVES: [OnceLock<Mutex<CurveGpu>>; 4] = [OnceLock::new(), OnceLock::new(), OnceLock::new(), OnceLock::new()];

pub fn gpu(curve_id: i32) -> &'static Mutex<CurveGpu> {
    CURVES[curve_id as usize].get_or_init(|| {
        let device = std::env::var("VIMZ_GPU_DEVICE").ok().and_then(|v| v.parse().ok()).unwrap_or(0);
        let ctx = sys::Context::new(curve_id, device).expect("vimz_ctx_create");
        Mutex::new(CurveGpu { ctx, keys: HashMap::new() })
    })
}

/// `&[G::Scalar]` / `&[G::PreprocessedGroupElement]` are `[u64; 4]` Montgomery limbs (halo2curves 0.1.0 / pasta_curves 0.5.1):
/// reinterpreted in place, no conversion -- the convention of pasta-msm's `mult_pippenger_pallas`.
#[inline]
unsafe fn as_scalars<T>(v: &[T]) -> &[Scalar] {
    debug_assert_eq!(std::mem::size_of::<T>(), 32);
    std::slice::from_raw_parts(v.as_ptr() as *const Scalar, v.len())
}
#[inline]
unsafe fn as_affines<T>(v: &[T]) -> &[Affine] {
    debug_assert_eq!(std::mem::size_of::<T>(), 64);
    std::slice::from_raw_parts(v.as_ptr() as *const Affine, v.len())
}

/// Body of `Group::vartime_multiscalar_mul(scalars, bases)`; `P` is the curve's projective point type (96 bytes, {x, y, z}).
pub fn msm<S, A, P: Copy>(curve_id: i32, scalars: &[S], bases: &[A]) -> P {
    assert!(bases.len() >= scalars.len());
    let mut g = gpu(curve_id).lock().unwrap();
    let key = (bases.as_ptr() as usize, bases.len());
    if !g.keys.contains_key(&key) {
        let ck = g.ctx.upload_key(unsafe { as_affines(bases) }).expect("vimz_ck_upload");
        g.keys.insert(key, ck);
    }
    let out: Point = g.ctx.commit(&g.keys[&key], unsafe { as_scalars(scalars) }).expect("vimz_msm");
    debug_assert_eq!(std::mem::size_of::<P>(), 96);
    unsafe { std::mem::transmute_copy::<Point, P>(&out) }
}

/// The resident fast path that replaces the body of `NIFS::prove` for one curve: r_U / r_W stay in HBM.
///
/// ```ignore
/// // src/nifs.rs, NIFS::prove
/// let (comm_w2, comm_t) = fold.step_begin(&W2.W, &U2.X);        // commit(W2) beside cross term + commit(T)
/// ro.absorb(pp_digest); U1.absorb_in_ro(&mut ro); U2.absorb_in_ro(&mut ro);
/// Commitment::<G>::from(comm_t).absorb_in_ro(&mut ro);
/// let r = ro.squeeze(NUM_CHALLENGE_BITS);
/// fold.step_end(&r);                                            // W, E, u, X, comm_W, comm_E folded on the GPU
/// ```
/// On the secondary curve, where prove_step commits at the end of step i and folds at the start of step i + 1, the two halves
/// are `commit_fresh` (inside r1cs_instance_and_witness) and `cross_begin` (inside NIFS::prove).
pub struct ResidentFold {
    pub acc: sys::Accumulator,
}

impl ResidentFold {
    pub fn step_begin<S>(&mut self, w2: &[S], x2: &[S]) -> (Point, Point) {
        self.acc.step_begin(unsafe { as_scalars(w2) }, unsafe { as_scalars(x2) }).expect("vimz_acc_step_begin")
    }
    pub fn step_end<S>(&mut self, r: &S) {
        self.acc.step_end(unsafe { &*(r as *const S as *const Scalar) }).expect("vimz_acc_step_end")
    }
}
