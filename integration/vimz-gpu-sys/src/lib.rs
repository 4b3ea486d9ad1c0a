//! FFI declarations + thin safe wrappers for `libvimz_gpu.so` (include/vimz_gpu.h).
//!
//! Layout contract: `Scalar`, `Affine`, `Point` below are the in-memory representations of the
//! halo2curves 0.1.0 / pasta_curves 0.5.1 types that nova-snark 0.23.0 instantiates
//! (`[u64; 4]` little-endian Montgomery limbs; affine identity = (0, 0); Jacobian identity Z = 0),
//! so `&[pallas::Scalar]` / `&[pallas::Affine]` are passed by pointer without conversion.
//!
//! Used from the patched provider: `vartime_multiscalar_mul`,
//! `R1CSShape::{multiply_vec, commit_T}`, `RelaxedR1CSWitness::fold`, and the resident accumulator that
//! replaces the body of `NIFS::prove` (call site: vimz/src/nova_snark_backend/folding.rs:35).
#![allow(non_camel_case_types)]

use std::ffi::CStr;
use std::os::raw::{c_char, c_int, c_long, c_void};

pub const VIMZ_PALLAS: c_int = 0;
pub const VIMZ_VESTA: c_int = 1;
pub const VIMZ_BN254: c_int = 2;
pub const VIMZ_GRUMPKIN: c_int = 3;

pub const VIMZ_ERR_LENGTH: c_int = -3; // -> NovaError::InvalidWitnessLength
pub const VIMZ_ERR_INDEX: c_int = -5; // -> NovaError::InvalidIndex

#[repr(C)]
#[derive(Clone, Copy, Default, Debug, PartialEq, Eq)]
pub struct Scalar(pub [u64; 4]);
#[repr(C)]
#[derive(Clone, Copy, Default, Debug, PartialEq, Eq)]
pub struct Affine { pub x: [u64; 4], pub y: [u64; 4] }
#[repr(C)]
#[derive(Clone, Copy, Default, Debug, PartialEq, Eq)]
pub struct Point { pub x: [u64; 4], pub y: [u64; 4], pub z: [u64; 4] }

#[repr(C)] pub struct vimz_ctx { _p: [u8; 0] }
#[repr(C)] pub struct vimz_ck { _p: [u8; 0] }
#[repr(C)] pub struct vimz_shape { _p: [u8; 0] }
#[repr(C)] pub struct vimz_acc { _p: [u8; 0] }
#[repr(C)] pub struct vimz_comm { _p: [u8; 0] }

extern "C" {
    pub fn vimz_last_error() -> *const c_char;
    /// page-locked host memory for witness buffers (a pageable Vec<Scalar> is staged by the driver at ~10 GB/s)
    pub fn vimz_host_alloc(bytes: usize) -> *mut c_void;
    pub fn vimz_host_free(p: *mut c_void);
    pub fn vimz_ctx_create(curve_id: c_int, device: c_int, out: *mut *mut vimz_ctx) -> c_int;
    pub fn vimz_ctx_destroy(ctx: *mut vimz_ctx);
    pub fn vimz_ctx_set_option(ctx: *mut vimz_ctx, key: *const c_char, value: c_long) -> c_int;
    pub fn vimz_ck_upload(ctx: *mut vimz_ctx, bases: *const Affine, n: usize, out: *mut *mut vimz_ck) -> c_int;
    pub fn vimz_ck_destroy(ck: *mut vimz_ck);
    pub fn vimz_msm(ctx: *mut vimz_ctx, ck: *const vimz_ck, scalars: *const Scalar, n: usize, out: *mut Point) -> c_int;
    pub fn vimz_shape_upload(ctx: *mut vimz_ctx, num_cons: usize, num_vars: usize, num_io: usize,
        row_a: *const u32, col_a: *const u32, val_a: *const Scalar, nnz_a: usize,
        row_b: *const u32, col_b: *const u32, val_b: *const Scalar, nnz_b: usize,
        row_c: *const u32, col_c: *const u32, val_c: *const Scalar, nnz_c: usize,
        out: *mut *mut vimz_shape) -> c_int;
    pub fn vimz_shape_destroy(s: *mut vimz_shape);
    pub fn vimz_multiply_vec(ctx: *mut vimz_ctx, s: *const vimz_shape, z: *const Scalar, z_len: usize,
        az: *mut Scalar, bz: *mut Scalar, cz: *mut Scalar) -> c_int;
    pub fn vimz_commit_T(ctx: *mut vimz_ctx, s: *const vimz_shape, ck: *const vimz_ck,
        w1: *const Scalar, u1: *const Scalar, x1: *const Scalar, w2: *const Scalar, x2: *const Scalar,
        t_out: *mut Scalar, comm_t: *mut Point) -> c_int;
    pub fn vimz_fold_witness(ctx: *mut vimz_ctx, r: *const Scalar, w1: *const Scalar, w2: *const Scalar, n: usize,
        e1: *const Scalar, t: *const Scalar, m: usize, w_out: *mut Scalar, e_out: *mut Scalar) -> c_int;
    pub fn vimz_acc_init(ctx: *mut vimz_ctx, s: *const vimz_shape, ck: *const vimz_ck, out: *mut *mut vimz_acc) -> c_int;
    /// Row-range shard of one fold across GPUs (include/vimz_gpu.h): partial commitments per rank.
    pub fn vimz_acc_init_sharded(ctx: *mut vimz_ctx, s_rows: *const vimz_shape, ck_rows: *const vimz_ck, ck_vars: *const vimz_ck,
        var_first: usize, var_count: usize, out: *mut *mut vimz_acc) -> c_int;
    pub fn vimz_acc_step_begin_dev_async(acc: *mut vimz_acc, d_w2: *const core::ffi::c_void, x2: *const Scalar,
        d_partials: *mut *mut core::ffi::c_void) -> c_int;
    pub fn vimz_acc_step_combine_dev(acc: *mut vimz_acc, d_gathered: *const core::ffi::c_void, world: usize,
        comm_w2: *mut Point, comm_t: *mut Point) -> c_int;
    pub fn vimz_point_sum(ctx: *mut vimz_ctx, pts: *const Point, k: usize, out: *mut Point) -> c_int;
    pub fn vimz_acc_load(acc: *mut vimz_acc, w: *const Scalar, e: *const Scalar, u: *const Scalar, x: *const Scalar,
        comm_w: *const Point, comm_e: *const Point) -> c_int;
    pub fn vimz_acc_step_begin(acc: *mut vimz_acc, w2: *const Scalar, x2: *const Scalar, comm_w2: *mut Point, comm_t: *mut Point) -> c_int;
    pub fn vimz_acc_step_end(acc: *mut vimz_acc, r: *const Scalar) -> c_int;
    pub fn vimz_acc_download(acc: *mut vimz_acc, w: *mut Scalar, e: *mut Scalar, u: *mut Scalar, x: *mut Scalar,
        comm_w: *mut Point, comm_e: *mut Point) -> c_int;
    pub fn vimz_acc_last_T(acc: *mut vimz_acc, t: *mut Scalar) -> c_int;
    pub fn vimz_acc_destroy(acc: *mut vimz_acc);
    pub fn vimz_point_to_affine(ctx: *mut vimz_ctx, p: *const Point, out: *mut Affine) -> c_int;
    pub fn vimz_point_scale_add(ctx: *mut vimz_ctx, a: *const Point, r: *const Scalar, b: *const Point, out: *mut Point) -> c_int;
    #[allow(dead_code)]
    fn vimz_ctx_sync(ctx: *mut vimz_ctx) -> c_int;
    // prove_step order on the secondary curve: commit at the end of step i, fold at the start of step i + 1
    pub fn vimz_acc_commit_fresh(acc: *mut vimz_acc, w2: *const Scalar, x2: *const Scalar, comm_w2: *mut Point) -> c_int;
    pub fn vimz_acc_cross_begin(acc: *mut vimz_acc, comm_t: *mut Point) -> c_int;
    pub fn vimz_acc_fresh_witness(acc: *mut vimz_acc, w2: *mut Scalar, x2: *mut Scalar) -> c_int;
    pub fn vimz_acc_reset(acc: *mut vimz_acc) -> c_int;
    // streaming upload of the fold-independent part of a witness
    pub fn vimz_acc_stage_fresh(acc: *mut vimz_acc, w2_part: *const Scalar, first: usize, count: usize) -> c_int;
    pub fn vimz_acc_step_begin_async(acc: *mut vimz_acc, w2: *const Scalar, x2: *const Scalar) -> c_int;
    pub fn vimz_acc_step_wait(acc: *mut vimz_acc, comm_w2: *mut Point, comm_t: *mut Point) -> c_int;
    pub fn vimz_acc_step_begin_staged(acc: *mut vimz_acc, w2_rest: *const Scalar, first: usize, count: usize, x2: *const Scalar,
        comm_w2: *mut Point, comm_t: *mut Point) -> c_int;
    // several GPUs of one node: NCCL inside the library (bound with dlopen), one process / thread per GPU
    pub fn vimz_comm_unique_id(id: *mut u8 /* [u8; 128] */) -> c_int;
    pub fn vimz_comm_create(device: c_int, id: *const u8, rank: c_int, world: c_int, out: *mut *mut vimz_comm) -> c_int;
    pub fn vimz_comm_destroy(comm: *mut vimz_comm);
    pub fn vimz_comm_broadcast_dev(ctx: *mut vimz_ctx, comm: *mut vimz_comm, d_buf: *mut core::ffi::c_void, bytes: usize, root: c_int) -> c_int;
    pub fn vimz_msm_sharded_dev(ctx: *mut vimz_ctx, comm: *mut vimz_comm, ck: *const vimz_ck, first: usize,
        d_scalars: *const core::ffi::c_void, n: usize, out: *mut Point) -> c_int;
    pub fn vimz_acc_step_begin_sharded_dev(acc: *mut vimz_acc, comm: *mut vimz_comm, d_w2: *const core::ffi::c_void, x2: *const Scalar,
        comm_w2: *mut Point, comm_t: *mut Point) -> c_int;
    pub fn vimz_acc_step_begin_sharded(acc: *mut vimz_acc, comm: *mut vimz_comm, w2: *const Scalar /* NULL off the root */, root: c_int,
        x2: *const Scalar, comm_w2: *mut Point, comm_t: *mut Point) -> c_int;
}

#[derive(Debug)]
pub struct GpuError { pub code: i32, pub message: String }

fn check(rc: c_int) -> Result<(), GpuError> {
    if rc == 0 { return Ok(()); }
    let message = unsafe { CStr::from_ptr(vimz_last_error()) }.to_string_lossy().into_owned();
    Err(GpuError { code: rc, message })
}

/// One curve on one GPU.  `CommitmentEngineTrait::commit` is infallible in nova-snark, so the provider
/// `expect`s these results exactly like the `assert!`s it replaces.
pub struct Context { raw: *mut vimz_ctx }
unsafe impl Send for Context {}

impl Context {
    pub fn new(curve_id: i32, device: i32) -> Result<Self, GpuError> {
        let mut raw = std::ptr::null_mut();
        check(unsafe { vimz_ctx_create(curve_id, device, &mut raw) })?;
        Ok(Self { raw })
    }
    pub fn upload_key(&self, bases: &[Affine]) -> Result<CommitmentKey, GpuError> {
        let mut ck = std::ptr::null_mut();
        check(unsafe { vimz_ck_upload(self.raw, bases.as_ptr(), bases.len(), &mut ck) })?;
        Ok(CommitmentKey { raw: ck })
    }
    /// `CE::commit(ck, v)`
    pub fn commit(&self, ck: &CommitmentKey, v: &[Scalar]) -> Result<Point, GpuError> {
        let mut out = Point::default();
        check(unsafe { vimz_msm(self.raw, ck.raw, v.as_ptr(), v.len(), &mut out) })?;
        Ok(out)
    }
    pub fn raw(&self) -> *mut vimz_ctx { self.raw }
    #[allow(dead_code)]
    pub fn raw_void(&self) -> *mut c_void { self.raw as *mut c_void }
}
impl Drop for Context { fn drop(&mut self) { unsafe { vimz_ctx_destroy(self.raw) } } }

pub struct CommitmentKey { raw: *mut vimz_ck }
unsafe impl Send for CommitmentKey {}
impl CommitmentKey { pub fn raw(&self) -> *const vimz_ck { self.raw } }
impl Drop for CommitmentKey { fn drop(&mut self) { unsafe { vimz_ck_destroy(self.raw) } } }

/// Device-resident running instance: `step_begin` = commit(W2) || cross term + commit(T); the caller squeezes `r`
/// from its (untouched) Poseidon RO; `step_end` folds W, E, u, X and both commitments on the GPU.
pub struct Accumulator { raw: *mut vimz_acc }
unsafe impl Send for Accumulator {}
impl Accumulator {
    /// # Safety: `shape` and `ck` must outlive the accumulator and belong to `ctx`.
    pub unsafe fn new(ctx: &Context, shape: *const vimz_shape, ck: &CommitmentKey) -> Result<Self, GpuError> {
        let mut raw = std::ptr::null_mut();
        check(vimz_acc_init(ctx.raw, shape, ck.raw, &mut raw))?;
        Ok(Self { raw })
    }
    pub fn step_begin(&mut self, w2: &[Scalar], x2: &[Scalar]) -> Result<(Point, Point), GpuError> {
        let (mut cw, mut ct) = (Point::default(), Point::default());
        check(unsafe { vimz_acc_step_begin(self.raw, w2.as_ptr(), x2.as_ptr(), &mut cw, &mut ct) })?;
        Ok((cw, ct))
    }
    pub fn step_end(&mut self, r: &Scalar) -> Result<(), GpuError> { check(unsafe { vimz_acc_step_end(self.raw, r) }) }
    /// r1cs_instance_and_witness on the secondary curve: commit now, fold at the start of the next prove_step.
    pub fn commit_fresh(&mut self, w2: &[Scalar], x2: &[Scalar]) -> Result<Point, GpuError> {
        let mut cw = Point::default();
        check(unsafe { vimz_acc_commit_fresh(self.raw, w2.as_ptr(), x2.as_ptr(), &mut cw) })?;
        Ok(cw)
    }
    pub fn cross_begin(&mut self) -> Result<Point, GpuError> {
        let mut ct = Point::default();
        check(unsafe { vimz_acc_cross_begin(self.raw, &mut ct) })?;
        Ok(ct)
    }
    /// Hand over the rows of the witness that do not depend on the previous fold while the other curve is being folded.
    /// # Safety: `part` must stay alive and unchanged until the next `step_begin_staged` returns.
    pub unsafe fn stage_fresh(&mut self, part: &[Scalar], first: usize) -> Result<(), GpuError> {
        check(vimz_acc_stage_fresh(self.raw, part.as_ptr(), first, part.len()))
    }
    pub fn step_begin_staged(&mut self, rest: &[Scalar], first: usize, x2: &[Scalar]) -> Result<(Point, Point), GpuError> {
        let (mut cw, mut ct) = (Point::default(), Point::default());
        check(unsafe { vimz_acc_step_begin_staged(self.raw, rest.as_ptr(), first, rest.len(), x2.as_ptr(), &mut cw, &mut ct) })?;
        Ok((cw, ct))
    }
    /// The same without waiting: the step is enqueued, `step_wait` collects the commitments (host sequencing of the two curves:
    /// the other curve's `step_end` is issued in between).
    /// # Safety: `rest` and `x2` must stay alive and unchanged until `step_wait` returns.
    pub unsafe fn step_begin_staged_async(&mut self, rest: &[Scalar], first: usize, x2: &[Scalar]) -> Result<(), GpuError> {
        check(vimz_acc_step_begin_staged(self.raw, rest.as_ptr(), first, rest.len(), x2.as_ptr(), std::ptr::null_mut(), std::ptr::null_mut()))
    }
    /// # Safety: `w2` and `x2` must stay alive and unchanged until `step_wait` returns.
    pub unsafe fn step_begin_async(&mut self, w2: &[Scalar], x2: &[Scalar]) -> Result<(), GpuError> {
        check(vimz_acc_step_begin_async(self.raw, w2.as_ptr(), x2.as_ptr()))
    }
    pub fn step_wait(&mut self) -> Result<(Point, Point), GpuError> {
        let (mut cw, mut ct) = (Point::default(), Point::default());
        check(unsafe { vimz_acc_step_wait(self.raw, &mut cw, &mut ct) })?;
        Ok((cw, ct))
    }
    pub fn raw(&self) -> *mut vimz_acc { self.raw }
}
impl Drop for Accumulator { fn drop(&mut self) { unsafe { vimz_acc_destroy(self.raw) } } }
