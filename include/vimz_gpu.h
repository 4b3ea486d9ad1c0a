/* vimz_gpu.h -- C ABI of the B200-native Nova fold kernels (libvimz_gpu.so).
 *
 * This is the drop-in boundary for the per-step NIFS fold that zero-savvy/vimz runs through
 * nova-snark 0.23.0 (hot loop entered at /root/reference/vimz/src/nova_snark_backend/folding.rs:35,
 * curve cycle selected at /root/reference/vimz/src/nova_snark_backend/mod.rs:19-20).  Each entry
 * point names the reference interface it replaces; nova-snark itself is an un-vendored dependency
 * (vimz/Cargo.toml:51, Cargo.lock:3577), so those are crate-relative paths.
 *
 * Conventions (identical to the in-memory layout of halo2curves 0.1.0 / pasta_curves 0.5.1):
 *   - field element  = 32 bytes, 4 x u64 little-endian limbs, MONTGOMERY form with R = 2^256;
 *   - affine point   = {x, y}, 64 bytes, identity encoded as (0, 0);
 *   - group element  = Jacobian {X, Y, Z}, 96 bytes, identity Z = 0 (returned as (0, R, 0));
 *   - every function returns 0 on success, a negative vimz_status otherwise; the message is
 *     available from vimz_last_error() (thread-local).  No entry point has a CPU fallback.
 *   - "host" pointers are borrowed for the duration of the call; "_dev" variants take device
 *     pointers (e.g. torch tensor .data_ptr()) that live on the context's device.
 *   - one context = one curve on one GPU with its own streams and workspaces.  Entry points may be called from
 *     any host thread: calls on the SAME context are serialised inside the library (a per-context lock held for
 *     the call -- CompressedSNARK::prove and RecursiveSNARK::verify commit from several rayon workers at once);
 *     use one context per curve, or one per worker if the commits should overlap on the GPU.
 */
#ifndef VIMZ_GPU_H
#define VIMZ_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { uint64_t l[4]; } vimz_fr;                 /* scalar-field element (Montgomery) */
typedef struct { uint64_t x[4], y[4]; } vimz_affine;       /* base-field affine point */
typedef struct { uint64_t x[4], y[4], z[4]; } vimz_point;  /* Jacobian group element */

typedef struct vimz_ctx vimz_ctx;     /* curve + device + stream + workspace */
typedef struct vimz_ck vimz_ck;       /* resident commitment key (window table in HBM) */
typedef struct vimz_shape vimz_shape; /* resident R1CSShape (CSR A, B, C) */
typedef struct vimz_acc vimz_acc;     /* resident running relaxed instance/witness */

enum vimz_curve { VIMZ_PALLAS = 0, VIMZ_VESTA = 1, VIMZ_BN254 = 2, VIMZ_GRUMPKIN = 3 };

enum vimz_status {
  VIMZ_OK = 0,
  VIMZ_ERR_CUDA = -1,     /* a CUDA runtime call or kernel failed */
  VIMZ_ERR_ARG = -2,      /* bad argument (null pointer, unknown curve, ...) */
  VIMZ_ERR_LENGTH = -3,   /* nova-snark's InvalidWitnessLength / ck shorter than the vector */
  VIMZ_ERR_NO_DEVICE = -4,/* no usable CUDA device: the library never computes on the CPU */
  VIMZ_ERR_INDEX = -5     /* R1CS entry out of range (nova-snark's InvalidIndex) */
};

const char* vimz_last_error(void);
int vimz_version(void);
int vimz_device_count(void);
/* Page-locked host memory for witness buffers: a witness handed over from ordinary (pageable) memory -- a Rust Vec<Scalar> -- is
 * staged by the driver at ~10 GB/s (0.4 ms for a grayscale-HD witness, measured: e2e.pageable_value of bench.py), from memory
 * obtained here it travels at PCIe speed and may be read asynchronously (vimz_acc_stage_fresh, vimz_acc_step_begin_async).
 * A host keeps two such buffers per curve and lets its witness generator write into them.  Returns NULL on failure. */
void* vimz_host_alloc(size_t bytes);
void vimz_host_free(void* p);

/* ---- context -------------------------------------------------------------------------------- */
/* Replaces the choice of provider made by `type G1/G2` (mod.rs:19-20): one context per curve. */
int vimz_ctx_create(int curve_id, int device, vimz_ctx** out);
void vimz_ctx_destroy(vimz_ctx* ctx);
int vimz_ctx_sync(vimz_ctx* ctx);
/* Tunables: "msm_window" (c bits, 0 = auto), "msm_acc_blocks" (accumulation blocks per SM, 1..8), "msm_seg_min" (shortest accumulation segment, 1..4096),
 * "msm_direct_c" (digit width of that table, 4 .. 14, 0 = by key length: 10 up to 16 384 points, else 8; 64 B << (c - 1) per point and window),
 * "msm_direct_max" (keys uploaded afterwards with at most this many points keep ALL digit multiples resident -- 256 KB per point at c = 8 --
 * and commit without buckets; default 32768, 0 = always the bucket pipeline; a forced msm_window also selects buckets), "msm_defer_giants" (0/1, default 1: buckets cut into hundreds of segments are summed beside
 * the bucket reduction instead of in front of it), "stage_commit" (0/1, default 0, see vimz_acc_stage_fresh), "acc_order" (0/1, default 0: commit(T)'s accumulation
 * kernel starts when commit(W2)'s has finished instead of sharing the SMs with it), "bitrow_fold" (0/1, default 1: accumulators created afterwards
 * on shapes with many booleanity rows b*(b-1)=0 keep K_S = sum over those rows of (A z1)_i ck_i and commit T + [row] A z1 instead of T -- half
 * of those rows then insert nothing; same comm_T), "spin_wait" (0/1, default 1: step_begin polls its stream instead of a blocking wait), "cross_cache" (0/1, default 1:
 * accumulators created afterwards keep (Az1, Bz1, Cz1) of the running instance resident and fold them in step_end instead of
 * recomputing them in every step_begin), "aux_lane" (0/1), "profile" (0/1), "graph" (0/1: replay a fold step's launch
 * sequence as a CUDA graph, default 1).  Unknown keys -> VIMZ_ERR_ARG. */
int vimz_ctx_set_option(vimz_ctx* ctx, const char* key, long value);
/* Device-side phase timers, enabled with vimz_ctx_set_option(ctx, "profile", 1): accumulated CUDA-event
 * milliseconds and call counts for "msm_sort", "msm_accumulate" (kernel + combine), "msm_accumulate_kernel", "msm_accumulate_kernel_T" (the commit(T) launches alone; "msm_entries_T" their insertions),
 * "msm_reduce", "cross_term", "axpy", "spmv"; name "msm_entries" returns the number of bucket insertions in *calls.  Synchronises the stream. */
int vimz_ctx_profile(vimz_ctx* ctx, const char* name, double* ms, uint64_t* calls, int reset);
/* The context's CUDA stream (cudaStream_t as void*), so a caller can time on it with events. */
void* vimz_ctx_stream(vimz_ctx* ctx);
/* Number of kernels this context has launched so far (bench.py's "gpu_launches"). */
uint64_t vimz_ctx_launch_count(vimz_ctx* ctx);

/* ---- commitment key / MSM -------------------------------------------------------------------- */
/* Replaces CommitmentEngineTrait::setup's product `CommitmentKey{ck: Vec<Affine>}` being read by
 * every commit ([EXT nova-snark] src/provider/pedersen.rs): the bases are uploaded ONCE and expanded
 * on the GPU into the window table {2^(c*j) * ck_i} that stays in HBM for all fold steps. */
int vimz_ck_upload(vimz_ctx* ctx, const vimz_affine* bases, size_t n, vimz_ck** out);
int vimz_ck_upload_dev(vimz_ctx* ctx, const void* d_bases, size_t n, vimz_ck** out);
void vimz_ck_destroy(vimz_ck* ck);
size_t vimz_ck_len(const vimz_ck* ck);
int vimz_ck_window_bits(const vimz_ck* ck);
int vimz_ck_num_windows(const vimz_ck* ck);

/* Replaces CommitmentEngineTrait::commit(ck, v) == Group::vartime_multiscalar_mul(v, ck[..v.len()])
 * ([EXT nova-snark] src/provider/pedersen.rs, src/provider/{pasta,bn256_grumpkin,mod}.rs;
 * C-ABI precedent: pasta-msm `mult_pippenger_pallas`).  n > len(ck) -> VIMZ_ERR_LENGTH. */
int vimz_msm(vimz_ctx* ctx, const vimz_ck* ck, const vimz_fr* scalars, size_t n, vimz_point* out);
int vimz_msm_dev(vimz_ctx* ctx, const vimz_ck* ck, const void* d_scalars, size_t n, vimz_point* out);
/* Point-range shard: sum_{i<n} scalars[i] * ck[first + i] (multi-GPU MSM). */
int vimz_msm_range_dev(vimz_ctx* ctx, const vimz_ck* ck, size_t first, const void* d_scalars, size_t n, vimz_point* out);
/* Enqueue only: the Jacobian result is written to d_out (96 bytes of device memory) on the context's
 * stream; no host synchronisation.  Used to time the kernels alone and by the fused step. */
int vimz_msm_async_dev(vimz_ctx* ctx, const vimz_ck* ck, size_t first, const void* d_scalars, size_t n, void* d_out);

/* Group helpers used on the boundary (all computed on the GPU).
 * vimz_point_sum: out = sum of k Jacobian points (combining per-GPU MSM shards).
 * vimz_point_to_affine: canonical affine (x, y) of a Jacobian point, identity -> (0,0)
 *   (what Commitment::to_coordinates / compress feed to the RO).
 * vimz_point_scale_add: out = a + r * b  (RelaxedR1CSInstance::fold's comm_W1 + r*comm_W2). */
int vimz_point_sum(vimz_ctx* ctx, const vimz_point* pts, size_t k, vimz_point* out);
int vimz_point_to_affine(vimz_ctx* ctx, const vimz_point* p, vimz_affine* out);
int vimz_point_scale_add(vimz_ctx* ctx, const vimz_point* a, const vimz_fr* r, const vimz_point* b, vimz_point* out);

/* ---- R1CS shape ------------------------------------------------------------------------------ */
/* Replaces R1CSShape::new ([EXT nova-snark] src/r1cs.rs): COO triples (row, col, val) per matrix in
 * constraint order, column space z = (W || u || X) of length num_vars + 1 + num_io.  Converted to
 * CSR once and kept resident. */
int vimz_shape_upload(vimz_ctx* ctx, size_t num_cons, size_t num_vars, size_t num_io,
                      const uint32_t* rowA, const uint32_t* colA, const vimz_fr* valA, size_t nnzA,
                      const uint32_t* rowB, const uint32_t* colB, const vimz_fr* valB, size_t nnzB,
                      const uint32_t* rowC, const uint32_t* colC, const vimz_fr* valC, size_t nnzC,
                      vimz_shape** out);
void vimz_shape_destroy(vimz_shape* s);

/* Replaces R1CSShape::multiply_vec(z) -> (Az, Bz, Cz).  z_len must equal num_vars + 1 + num_io,
 * else VIMZ_ERR_LENGTH (nova-snark: NovaError::InvalidWitnessLength). */
int vimz_multiply_vec(vimz_ctx* ctx, const vimz_shape* s, const vimz_fr* z, size_t z_len,
                      vimz_fr* Az, vimz_fr* Bz, vimz_fr* Cz);

/* Replaces R1CSShape::commit_T(ck, U1, W1, U2, W2) -> (T, comm_T):
 *   T = Az1 o Bz2 + Az2 o Bz1 - u1*Cz2 - u2*Cz1 (u2 = 1), comm_T = commit(ck, T).
 * T_out may be NULL. */
int vimz_commit_T(vimz_ctx* ctx, const vimz_shape* s, const vimz_ck* ck,
                  const vimz_fr* W1, const vimz_fr* u1, const vimz_fr* X1,
                  const vimz_fr* W2, const vimz_fr* X2,
                  vimz_fr* T_out, vimz_point* comm_T);

/* Replaces RelaxedR1CSWitness::fold: W = W1 + r*W2 (n), E = E1 + r*T (m). */
int vimz_fold_witness(vimz_ctx* ctx, const vimz_fr* r,
                      const vimz_fr* W1, const vimz_fr* W2, size_t n,
                      const vimz_fr* E1, const vimz_fr* T, size_t m,
                      vimz_fr* W_out, vimz_fr* E_out);

/* ---- device-resident fold (the fast path) ---------------------------------------------------- */
/* The running relaxed witness (W1, E1), instance scalars (u1, X1) and their commitments stay in
 * HBM between steps; one step = NIFS::prove's data-parallel body ([EXT nova-snark] src/nifs.rs):
 *   step_begin: upload W2/X2 -> comm_W2 = commit(ck, W2) (the r1cs_instance_and_witness MSM),
 *               (Az,Bz,Cz)(z1), (Az,Bz,Cz)(z2), T, comm_T = commit(ck, T); returns both commitments.
 *   [host: RO absorbs comm_T, squeezes r -- untouched Poseidon RO]
 *   step_end:   W1 += r*W2, E1 += r*T, u1 += r, X1 += r*X2, comm_W1 += r*comm_W2, comm_E1 += r*comm_T.
 * vimz_acc_init starts from the default (all-zero, u = 0) relaxed instance like
 * RelaxedR1CSWitness::default / RelaxedR1CSInstance::default; vimz_acc_load starts from given values. */
int vimz_acc_init(vimz_ctx* ctx, const vimz_shape* s, const vimz_ck* ck, vimz_acc** out);
/* Row-range shard of one fold across GPUs: this rank's accumulator holds the constraint
 * rows [row_first, row_first + m_local) of A, B, C (`s_rows`: num_cons = m_local, same column space), the matching
 * slice of E / T, and commits with two resident key shards: ck_rows = ck[row_first ..) for T / E and
 * ck_vars = ck[var_first .. var_first + var_count) for its share of commit(W2).  W stays replicated (every rank
 * needs the whole z).  step_begin returns this rank's PARTIAL (comm_W2, comm_T); the caller all-gathers and adds
 * them (vimz_point_sum) before the RO; the running commitments kept here are partial sums as well (the fold
 * comm + r*comm' is linear), so vimz_acc_download returns shard values to be added across ranks. */
int vimz_acc_init_sharded(vimz_ctx* ctx, const vimz_shape* s_rows, const vimz_ck* ck_rows, const vimz_ck* ck_vars,
                          size_t var_first, size_t var_count, vimz_acc** out);
/* Sharded step without host round trips: step_begin_dev_async only ENQUEUES the step on the context stream and returns
 * the device address of this rank's partial (comm_W2, comm_T) pair (2 x 96 bytes, valid until the next step_begin); the
 * caller all-gathers those 192 bytes over NCCL on the same stream (vimz_ctx_stream) into d_gathered[world][2] and
 * step_combine_dev adds them on the GPU, copies the two full commitments back and synchronises once. */
int vimz_acc_step_begin_dev_async(vimz_acc* acc, const void* d_W2, const vimz_fr* X2, void** d_partials);
int vimz_acc_step_combine_dev(vimz_acc* acc, const void* d_gathered, size_t world, vimz_point* comm_W2, vimz_point* comm_T);
/* Back to the default relaxed instance (a new proof over the same shape and key). */
int vimz_acc_reset(vimz_acc* acc);
int vimz_acc_load(vimz_acc* acc, const vimz_fr* W, const vimz_fr* E, const vimz_fr* u, const vimz_fr* X,
                  const vimz_point* comm_W, const vimz_point* comm_E);
int vimz_acc_step_begin(vimz_acc* acc, const vimz_fr* W2, const vimz_fr* X2, vimz_point* comm_W2, vimz_point* comm_T);
int vimz_acc_step_begin_dev(vimz_acc* acc, const void* d_W2, const vimz_fr* X2, vimz_point* comm_W2, vimz_point* comm_T);
/* The two halves of step_begin as separate calls, for the SECONDARY curve of RecursiveSNARK::prove_step ([EXT nova-snark]
 * src/lib.rs): its fresh witness is committed at the end of step i (r1cs_instance_and_witness -> commit_fresh, which also
 * stages W2 / X2 in the accumulator) and folded at the start of step i+1 (NIFS::prove -> cross_begin: T and comm_T);
 * vimz_acc_step_end follows as usual.  vimz_acc_fresh_witness reads the staged (W2, X2) back (l_w_secondary, checked by
 * RecursiveSNARK::verify with is_sat before it is ever folded). */
int vimz_acc_commit_fresh(vimz_acc* acc, const vimz_fr* W2, const vimz_fr* X2, vimz_point* comm_W2);
int vimz_acc_cross_begin(vimz_acc* acc, vimz_point* comm_T);
int vimz_acc_fresh_witness(vimz_acc* acc, vimz_fr* W2, vimz_fr* X2);
/* Streaming upload of the fresh witness: vimz_acc_stage_fresh enqueues an H2D copy of W2[first .. first + count) behind the
 * previous step_end and returns at once (the host buffer must stay valid until the next step_begin* returns); the part of a
 * witness that does not depend on the previous fold (the Circom step circuit's variables) can so travel while the other curve
 * is being folded.  vimz_acc_step_begin_staged uploads the remaining range and runs the step; with comm_W2 = comm_T = NULL it only
 * enqueues the step and vimz_acc_step_wait collects the commitments.  The witness pointers of stage_fresh, step_begin_staged and
 * step_begin_async may also be DEVICE memory (unified addressing tells): a resident witness is then copied on the GPU.  With the
 * context option "stage_commit" = 1 a staged prefix / suffix of at least 1024 variables is also COMMITTED at once, on a lane of
 * its own beside whatever the GPU does for the other curve; the step then commits only the rest and adds the two (same comm_W2;
 * the staged rows must not change afterwards -- re-uploading them in step_begin_staged voids the early commitment). */
int vimz_acc_stage_fresh(vimz_acc* acc, const vimz_fr* W2_part, size_t first, size_t count);
/* vimz_acc_step_begin in two halves: _async copies W2 / X2 and enqueues the step (the host buffers must stay valid until _wait
 * returns), _wait blocks for the commitments.  Host-to-device copies of different accumulators share one copy engine in issue
 * order: a host that stages the NEXT primary witness while the secondary curve folds issues secondary _async, primary
 * stage_fresh, secondary _wait -- the small secondary witness goes first and the large staged copy hides behind its kernels. */
int vimz_acc_step_begin_async(vimz_acc* acc, const vimz_fr* W2, const vimz_fr* X2);
int vimz_acc_step_wait(vimz_acc* acc, vimz_point* comm_W2, vimz_point* comm_T);
int vimz_acc_step_begin_staged(vimz_acc* acc, const vimz_fr* W2_rest, size_t first, size_t count, const vimz_fr* X2, vimz_point* comm_W2,
                               vimz_point* comm_T);
int vimz_acc_step_end(vimz_acc* acc, const vimz_fr* r);
int vimz_acc_download(vimz_acc* acc, vimz_fr* W, vimz_fr* E, vimz_fr* u, vimz_fr* X, vimz_point* comm_W, vimz_point* comm_E);
/* T of the last step_begin (m elements), for callers that keep nova-snark's (T, comm_T) pair. */
int vimz_acc_last_T(vimz_acc* acc, vimz_fr* T);
void vimz_acc_destroy(vimz_acc* acc);

/* ---- several GPUs of one node ------------------------------------------ */
/* One process (or thread) per GPU.  A communicator wraps an NCCL communicator bound at run time (dlopen of libnccl.so.2 on
 * first use -- single-GPU users never load NCCL); rank 0 makes the 128-byte id and hands it to the other ranks over any
 * channel the host has.  Collectives are enqueued on the stream of the context / accumulator they are called with, so the
 * exchange is ordered with the kernels and the host waits once per call.  The reference has no counterpart: its
 * parallelism is rayon inside one process; these entry points are what a multi-GPU build of the
 * provider selected at /root/reference/vimz/src/nova_snark_backend/mod.rs:19-20 would call instead of vimz_msm /
 * vimz_acc_step_begin. */
typedef struct vimz_comm vimz_comm;
#define VIMZ_COMM_ID_BYTES 128
int vimz_comm_unique_id(uint8_t id[VIMZ_COMM_ID_BYTES]);
int vimz_comm_create(int device, const uint8_t id[VIMZ_COMM_ID_BYTES], int rank, int world, vimz_comm** out);
void vimz_comm_destroy(vimz_comm* comm);
int vimz_comm_rank(const vimz_comm* comm);
int vimz_comm_world(const vimz_comm* comm);
int vimz_comm_nccl_version(void);
/* NCCL broadcast of `bytes` of device memory from rank `root`, on the context's stream. */
int vimz_comm_broadcast_dev(vimz_ctx* ctx, vimz_comm* comm, void* d_buf, size_t bytes, int root);
/* commit(ck, v) with ck / v split by point range across the ranks: this rank passes its slice (first, d_scalars, n) of a key
 * shard it uploaded; the 96-byte partial sums are all-gathered and added on the GPU, every rank gets the full commitment. */
int vimz_msm_sharded_dev(vimz_ctx* ctx, vimz_comm* comm, const vimz_ck* ck, size_t first, const void* d_scalars, size_t n, vimz_point* out);
/* One fold step of a row-range shard (vimz_acc_init_sharded): enqueue, all-gather the partial (comm_W2, comm_T) pairs,
 * add them on the GPU, wait once; every rank returns the FULL commitments (and then derives the same challenge).
 * _dev: W2 is device memory on every rank.  Host variant: W2 is host memory on rank `root` (NULL elsewhere); the root
 * copies it to its GPU and NCCL broadcasts it into every rank's accumulator before the step runs. */
int vimz_acc_step_begin_sharded_dev(vimz_acc* acc, vimz_comm* comm, const void* d_W2, const vimz_fr* X2, vimz_point* comm_W2, vimz_point* comm_T);
int vimz_acc_step_begin_sharded(vimz_acc* acc, vimz_comm* comm, const vimz_fr* W2, int root, const vimz_fr* X2, vimz_point* comm_W2,
                                vimz_point* comm_T);

/* ---- test / bench utilities (device-side generators; not part of the reference interface) ---- */
/* bases[i] = (k0 + i*dk) * G written as n affine points to device memory d_out (64*n bytes). */
int vimz_gen_bases_dev(vimz_ctx* ctx, uint64_t k0, uint64_t dk, size_t n, void* d_out);
/* Element-wise Montgomery product out[i] = a[i]*b[i] in the curve's BASE (which=0) or SCALAR (which=1)
 * field, plus sum/difference: op 0 = mul, 1 = add, 2 = sub, 3 = square of a (b is ignored).  Exercises fp.cuh directly for
 * the parity tests. */
int vimz_field_op(vimz_ctx* ctx, int which, int op, const vimz_fr* a, const vimz_fr* b, size_t n, vimz_fr* out);

#ifdef __cplusplus
}
#endif
#endif /* VIMZ_GPU_H */
