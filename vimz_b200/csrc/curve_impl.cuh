// curve_impl.cuh -- host-side launch sequences, templated on the curve; each curve_<name>.cu
// instantiates one CurveVTable so the four curves compile in parallel translation units.
#pragma once
#include <algorithm>
#include <cstring>
#include "common.cuh"
#include "msm.cuh"
#include "r1cs.cuh"

#ifndef VIMZ_BIG_BLOCKS_PER_SM
#define VIMZ_BIG_BLOCKS_PER_SM 1  // blocks of the giant-bucket role of k_msm_combine_all per SM
#endif

#ifndef VIMZ_TAIL_PER_QUAD
#define VIMZ_TAIL_PER_QUAD 4  // chunk sums a quad of k_reduce_tail adds serially before the block tree (A/B at c = 15: 2 -> +9 us, 8 -> +40 us)
#endif
namespace vimz {

struct CurveVTable {
  int id;
  const char* name;
  uint32_t scalar_modulus[8];
  uint32_t scalar_one_mont[8];
  int (*init_device)(vimz_ctx*);  // per-device kernel attributes, once per context
  int (*precompute)(vimz_ctx*, const void* d_bases, size_t n, int c, int nwin, void* table);
  // dtable[((j*n + i) << (c-1)) + k-1] = k * table[j][i]  (direct-table keys)
  int (*precompute_direct)(vimz_ctx*, const void* table, size_t n, int c, int nwin, void* dtable);
  // lane 0 = context stream + main workspace, lane 1 = aux stream + second workspace (runs concurrently), lane 2 = early stream +
  // third workspace (early commitment of a staged witness range, beside another context's step)
  // (MsmWorkspace::host_out / sub_jac / sub_event, set by the caller around one call, reach the final kernel)
  // counted = true: the bucket histogram of the scalars is already in the lane's `counts` buffer (fused cross term)
  int (*msm)(vimz_ctx*, int lane, const vimz_ck*, size_t first, const void* d_scalars, size_t n, void* d_out, bool counted);
  int (*msm_digits)(vimz_ctx*, int lane, const vimz_ck*, const void* d_scalars, size_t n, bool* done);
  int (*point_sum)(vimz_ctx*, const void* d_pts, size_t k, void* d_out);
  int (*point_sum_batch)(vimz_ctx*, const void* d_pts, size_t k, size_t sets, void* d_out);
  // d_out = d_a + d_b (Jacobian), also written to `host_out` (mapped page-locked host memory) when given
  int (*point_add2)(vimz_ctx*, cudaStream_t, const void* d_a, const void* d_b, void* d_out, void* host_out);
  int (*point_to_affine)(vimz_ctx*, const void* d_pt, void* d_out);
  int (*point_scale_add)(vimz_ctx*, cudaStream_t, const void* d_a, const void* d_r, const void* d_b, void* d_out, int count);
  int (*point_scale_add_val)(vimz_ctx*, cudaStream_t, const void* d_a, const vimz_fr* r, const void* d_b, void* d_out, int count);
  int (*gen_bases)(vimz_ctx*, uint64_t k0, uint64_t dk, size_t n, void* d_out);
  int (*spmv3)(vimz_ctx*, const vimz_shape*, const void* d_W, const void* d_tail, void* d_Az, void* d_Bz, void* d_Cz);
  // fuse_ck != nullptr: also histogram T's digits for the commit(fuse_ck, T) that follows on lane 0
  // cache1 / cache2 != nullptr (resident accumulator): (Az1, Bz1, Cz1) are READ from cache1[3][m] instead of being
  // recomputed and (Az2, Bz2, Cz2) are written to cache2[3][m] for the fold in step_end
  // rowflag != nullptr: the digits recoded for fuse_ck are those of T + [rowflag] Az1 (the caller subtracts K_S from the commitment)
  int (*cross_term)(vimz_ctx*, const vimz_shape*, const void* d_W1, const void* d_tail1, const void* d_W2, const void* d_tail2, void* d_T,
                    const vimz_ck* fuse_ck, const void* cache1, void* cache2, const uint8_t* rowflag);
  int (*mask_rows)(vimz_ctx*, cudaStream_t, const void* d_v, const uint8_t* rowflag, size_t m, void* d_out);
  // d_out = sum over the booleanity rows i of vals[bitcol[i]] * ck_i (bases = row 0 of the key's window table; vals = the fresh W2);
  // scratch: masked_sum_scratch_bytes(), zeroed once
  int (*masked_base_sum)(vimz_ctx*, cudaStream_t, const void* d_vals, const uint32_t* bitcol, size_t m, const vimz_ck*, void* scratch, void* d_out);
  // d_parts[j] = 2^(32 j) * (Jacobian) d_pt as XYZZ records, j < SCALE_PARTS; then d_out = d_a + r * d_pt from those parts
  int (*point_pow2_parts)(vimz_ctx*, cudaStream_t, const void* d_pt, void* d_parts);
  int (*point_scale_add_parts)(vimz_ctx*, cudaStream_t, const void* d_a, const vimz_fr* r, const void* d_parts, void* d_out);
  int (*axpy)(vimz_ctx*, const void* d_a, const void* d_b, const vimz_fr* r, size_t len, void* d_out);
  // up to AXPY_MAX_SEGS in-place folds a_k += r * b_k in one launch (witness fold W, E, the (u, X) tail, cached products)
  int (*axpyn)(vimz_ctx*, const vimz::AxpySeg* segs, int count, const vimz_fr* r);
  int (*field_op)(vimz_ctx*, int which, int op, const void* d_a, const void* d_b, size_t n, void* d_out);
};

const CurveVTable* curve_vtable(int curve_id);

#define VIMZ_LAUNCH_CHECK(ctx)                 \
  do {                                         \
    (ctx)->launches++;                         \
    VIMZ_CUDA(cudaGetLastError());             \
  } while (0)

inline uint32_t ceil_div(size_t a, size_t b) { return (uint32_t)((a + b - 1) / b); }

constexpr size_t MATVEC_SMEM = (size_t)CROSS_CHUNK_NNZ * (8 + 32);  // index stream + product slots of one chunk

// Per-DEVICE kernel attributes (cudaFuncSetAttribute applies to the current device only): called by vimz_ctx_create
// for every context, so a process driving several GPUs gets the > 48 KB shared-memory opt-in on each of them.
template <class C>
int impl_init_device(vimz_ctx*) {
  VIMZ_CUDA(cudaFuncSetAttribute(k_matvec_stream<typename C::Fs>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MATVEC_SMEM));
  return VIMZ_OK;
}

template <class C>
int impl_precompute(vimz_ctx* ctx, const void* d_bases, size_t n, int c, int nwin, void* table) {
  if (n == 0) return VIMZ_OK;
  k_precompute<C><<<ceil_div(n, 128), 128, 0, ctx->stream>>>(d_bases, (uint32_t)n, c, nwin, table);
  VIMZ_LAUNCH_CHECK(ctx);
  return VIMZ_OK;
}

template <class C>
int impl_precompute_direct(vimz_ctx* ctx, const void* table, size_t n, int c, int nwin, void* dtable) {
  if (n == 0) return VIMZ_OK;
  const size_t entries = n * (size_t)nwin;
  k_precompute_direct<C><<<ceil_div(entries, 128), 128, 0, ctx->stream>>>(table, (uint32_t)entries, c - 1, dtable);
  VIMZ_LAUNCH_CHECK(ctx);
  return VIMZ_OK;
}

// MSM over a direct-table key: digits (unless the cross term already wrote them) + one launch.
template <class C>
int impl_msm_direct(vimz_ctx* ctx, cudaStream_t st, MsmWorkspace& ws, const vimz_ck* ck, size_t first, const void* d_scalars, size_t n,
                    void* d_out, bool digits_ready) {
  const int c = ck->c, nwin = ck->nwin;
  const size_t E = std::max<size_t>(n * (size_t)nwin, 1);
  const uint32_t max_blocks = std::min<uint32_t>((uint32_t)ctx->sm_count * (uint32_t)ctx->opt_direct_bps, DIRECT_MAX_BLOCKS);
  const uint32_t blocks = std::max<uint32_t>(1, std::min<uint32_t>(ceil_div(E, 128 * 2), max_blocks));
  const uint32_t ngroups = ceil_div(blocks, DIRECT_GROUP);
  VIMZ_TRY(ws.digits.reserve(E * 4));
  if (ws.cls.cap < 64 * 4) {  // arrival counters: zeroed once, the last block of every launch leaves them zero again (no memset node per commit)
    VIMZ_TRY(ws.cls.reserve(64 * 4));
    VIMZ_CUDA(cudaMemsetAsync(ws.cls.ptr, 0, 64 * 4, st));
  }
  VIMZ_TRY(ws.partials.reserve(((size_t)blocks + ngroups) * 128));
  ws.last_M = 0;
  if (n > 0 && !digits_ready) {
    ProfScope prof_sort(ctx, PROF_MSM_SORT, st);
    const int grid_n = (int)std::min<size_t>(ceil_div(n, 256), (size_t)ctx->sm_count * 8);
    k_msm_digits<C><<<grid_n, 256, 0, st>>>(reinterpret_cast<const uint32_t*>(d_scalars), (uint32_t)n, c, nwin, nullptr,
                                            ws.digits.as<uint32_t>());
    VIMZ_LAUNCH_CHECK(ctx);
  }
  ProfScope prof_acc(ctx, PROF_MSM_ACCUMULATE, st);
  ProfScope prof_kernel(ctx, PROF_MSM_ACC_KERNEL, st);
  k_msm_direct<C><<<blocks, 128, 0, st>>>(ws.digits.as<uint32_t>(), (uint32_t)n, nwin, (uint32_t)ck->n, (uint32_t)first, c - 1, ck->dtable,
                                          ws.partials.ptr, ws.cls.as<uint32_t>(), d_out, ws.host_out);
  VIMZ_LAUNCH_CHECK(ctx);
  return VIMZ_OK;
}

// Geometry shared by the bucket pipeline and the fused cross term: accumulation threads, capacity of the giant lists,
// and the control block, which lives IN FRONT of the bucket histogram so one memset node clears both.
inline uint32_t msm_nthreads(const vimz_ctx* ctx) { return (uint32_t)ctx->sm_count * (uint32_t)ctx->opt_acc_blocks * 128; }
inline uint32_t msm_max_giants(const vimz_ctx* ctx) { return 2 * msm_nthreads(ctx) / COMBINE_MID + 2; }
inline uint32_t msm_ctrl_words(const vimz_ctx* ctx) { return (CTRL_GIANT_DONE + msm_max_giants(ctx) + 3) & ~3u; }  // keeps counts[] 16-byte aligned
// zero `nz` 16-byte words at `z`, and copy `nc` 16-byte words src -> dst (src: mapped page-locked host memory)
static __global__ void __launch_bounds__(256) k_zero_and_copy(uint4* __restrict__ z, uint32_t nz, const uint4* __restrict__ src, uint4* __restrict__ dst, uint32_t nc) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nc) dst[i] = src[i];
  if (i < nz) z[i] = make_uint4(0, 0, 0, 0);
}
inline int msm_zero_control(vimz_ctx* ctx, MsmWorkspace& ws, uint32_t M, cudaStream_t st) {
  const size_t bytes = ((size_t)msm_ctrl_words(ctx) + M) * 4;  // (a multiple of 16: msm_ctrl_words is a multiple of 4, M a power of two >= 4 or ...)
  VIMZ_TRY(ws.counts.reserve((bytes + 15) & ~(size_t)15));
  if (ws.pro_bytes) {  // the caller's small host -> device copy rides along (MsmWorkspace::pro_*): one node instead of two
    const uint32_t nz = (uint32_t)((bytes + 15) / 16), nc = (uint32_t)(ws.pro_bytes / 16);
    k_zero_and_copy<<<ceil_div(std::max(nz, nc), 256u), 256, 0, st>>>(reinterpret_cast<uint4*>(ws.counts.ptr), nz,
                                                                      reinterpret_cast<const uint4*>(ws.pro_src), reinterpret_cast<uint4*>(ws.pro_dst), nc);
    VIMZ_LAUNCH_CHECK(ctx);
    ws.pro_bytes = 0;  // consumed
    return VIMZ_OK;
  }
  VIMZ_CUDA(cudaMemsetAsync(ws.counts.ptr, 0, bytes, st));
  return VIMZ_OK;
}

// The first two nodes of a bucket-pipeline commit -- clear control block + histogram, recode the scalars -- by themselves, so that a
// caller can create them ahead of another lane's nodes; the commit proper follows as msm(..., counted = true).  Returns
// VIMZ_OK without doing anything for a direct-table key or an empty range (msm is then called with counted = false).
template <class C>
int impl_msm_digits(vimz_ctx* ctx, int lane, const vimz_ck* ck, const void* d_scalars, size_t n, bool* done) {
  *done = false;
  if (ck->dtable || n == 0) return VIMZ_OK;
  cudaStream_t st = lane == 0 ? ctx->stream : (lane == 1 ? ctx->aux : ctx->early);
  MsmWorkspace& ws = lane == 0 ? ctx->ws : (lane == 1 ? ctx->ws_aux : ctx->ws_early);
  const int c = ck->c, nwin = ck->nwin;
  const uint32_t M = 1u << (c - 1);
  VIMZ_TRY(ws.digits.reserve(std::max<size_t>(n * (size_t)nwin, 1) * 4));
  VIMZ_TRY(msm_zero_control(ctx, ws, M, st));
  ProfScope prof_sort(ctx, PROF_MSM_SORT, st);
  const int grid_n = (int)std::min<size_t>(ceil_div(n, 256), (size_t)ctx->sm_count * 8);
  k_msm_digits<C><<<grid_n, 256, 0, st>>>(reinterpret_cast<const uint32_t*>(d_scalars), (uint32_t)n, c, nwin,
                                          ws.counts.as<uint32_t>() + msm_ctrl_words(ctx), ws.digits.as<uint32_t>());
  VIMZ_LAUNCH_CHECK(ctx);
  *done = true;
  return VIMZ_OK;
}

template <class C>
int impl_msm(vimz_ctx* ctx, int lane, const vimz_ck* ck, size_t first, const void* d_scalars, size_t n, void* d_out, bool counted) {
  cudaStream_t st = lane == 0 ? ctx->stream : (lane == 1 ? ctx->aux : ctx->early);
  MsmWorkspace& ws = lane == 0 ? ctx->ws : (lane == 1 ? ctx->ws_aux : ctx->ws_early);
  if (ck->dtable) {
    if (ws.sub_jac) return set_error(VIMZ_ERR_ARG, "msm: a subtracted point is only supported on the bucket pipeline");
    return impl_msm_direct<C>(ctx, st, ws, ck, first, d_scalars, n, d_out, counted);
  }
  const int c = ck->c, nwin = ck->nwin;
  const uint32_t M = 1u << (c - 1);
  const size_t E = std::max<size_t>(n * (size_t)nwin, 1);
  // reduction geometry: chunks of K buckets (deeper chunks only when there are many buckets)
  const int K = (int)std::min<uint32_t>(M >= (1u << 18) ? 8 : 4, M);
  int logK = 0;
  while ((1 << logK) < K) logK++;
  const uint32_t T = M / K;
  int nb = 0;
  while ((1u << nb) < T) nb++;
  // second level: 32 quads per block, ~VIMZ_TAIL_PER_QUAD chunk sums per quad: the tree depth, not the work, sets the time
  // (G * (nb + 1) blocks of 128 threads must stay resident together -- 4 per SM -- or the last-arriving hand-offs wait for a
  // second wave: with 32 768 buckets and G = 64 the tail took 1.5 waves and ~50 us longer)
  const uint32_t g_fit = std::max<uint32_t>(((uint32_t)ctx->sm_count * 4) / (uint32_t)(nb + 2), 1);
  const int G = (int)std::min<uint32_t>(std::min<uint32_t>(std::max<uint32_t>(ceil_div(std::max<uint32_t>(T / 2, 1), 32 * VIMZ_TAIL_PER_QUAD), 1), 64), g_fit);
  // accumulation geometry: a fixed number of threads (4 resident warps per scheduler) share the E insertions
  const uint32_t nthreads = msm_nthreads(ctx);
  const uint32_t seg_min = (uint32_t)(lane == 1 && ctx->opt_seg_min_aux ? ctx->opt_seg_min_aux : ctx->opt_seg_min);
  // capacity bounds: a bucket cut into p pieces overlaps p segments and every segment boundary cuts at most one
  // bucket, so sum(pieces) <= 2 * nthreads; a giant has > COMBINE_MID pieces and ceil(p / GIANT_CHUNK) chunks
  const uint32_t max_giants = msm_max_giants(ctx);
  const uint32_t max_chunks = 2 * nthreads / GIANT_CHUNK + max_giants + 2;
  const uint32_t ctrl_words = msm_ctrl_words(ctx);

  VIMZ_TRY(ws.counts.reserve(((size_t)ctrl_words + M) * 4));
  VIMZ_TRY(ws.offsets.reserve(((size_t)M + 1) * 4));
  VIMZ_TRY(ws.cursor.reserve((size_t)M * 4));
  uint32_t scan_blocks = ceil_div(M, SCAN_THREADS * SCAN_ITEMS);
  VIMZ_TRY(ws.blocksums.reserve(((size_t)scan_blocks + 2) * 4));
  VIMZ_TRY(ws.sorted.reserve(E * 4));
  VIMZ_TRY(ws.digits.reserve(E * 4));
  VIMZ_TRY(ws.biglist.reserve(((size_t)3 * max_giants + (size_t)2 * max_chunks + (size_t)M + 4) * 4));
  VIMZ_TRY(ws.partials.reserve(((size_t)2 * nthreads + max_chunks) * 128));
  VIMZ_TRY(ws.buckets.reserve((size_t)M * 128));
  VIMZ_TRY(ws.chunkA.reserve((size_t)T * 128));
  VIMZ_TRY(ws.chunkL.reserve((size_t)T * 128));
  VIMZ_TRY(ws.bitsums.reserve((size_t)(nb + 1) * G * 128));
  VIMZ_TRY(ws.scaled.reserve((size_t)(nb + 1) * 128));
  const bool defer = ctx->opt_defer_giants;
  if (defer) VIMZ_TRY(ws.deferred.reserve((size_t)max_giants * 128));

  uint32_t* counts = ws.counts.as<uint32_t>() + ctrl_words;
  uint32_t* offsets = ws.offsets.as<uint32_t>();
  uint32_t* cursor = ws.cursor.as<uint32_t>();
  uint32_t* blocksums = ws.blocksums.as<uint32_t>();
  uint32_t* sorted = ws.sorted.as<uint32_t>();
  MsmCombine cb;
  cb.ctrl = ws.counts.as<uint32_t>();
  cb.giants = ws.biglist.as<uint32_t>();
  cb.chunk_rec = cb.giants + (size_t)3 * max_giants;
  cb.mids = cb.chunk_rec + (size_t)2 * max_chunks;
  cb.chunk_sums = ws.partials.as<char>() + (size_t)2 * nthreads * 128;
  cb.max_giants = max_giants;
  cb.max_chunks = max_chunks;

  ws.last_M = M;
  const int phase = ws.phase;  // 0: everything, 1: up to the accumulation, 2: from the combine on (MsmWorkspace::phase)
  if (phase != 2 && !counted) VIMZ_TRY(msm_zero_control(ctx, ws, M, st));  // (counted: the fused cross term cleared control block + histogram)

  const int grid_n = (int)std::min<size_t>(ceil_div(std::max<size_t>(n, 1), 256), (size_t)ctx->sm_count * 8);
  if (phase != 2) {
    ProfScope prof_sort(ctx, PROF_MSM_SORT, st);
    if (n > 0 && !counted) {
      k_msm_digits<C><<<grid_n, 256, 0, st>>>(reinterpret_cast<const uint32_t*>(d_scalars), (uint32_t)n, c, nwin, counts,
                                              ws.digits.as<uint32_t>());
      VIMZ_LAUNCH_CHECK(ctx);
    }
    if (M <= SCAN1_MAX_M) {  // every fold-step MSM: one single-block launch
      k_scan_single<<<1, SCAN1_THREADS, 0, st>>>(counts, M, offsets, cursor);
      VIMZ_LAUNCH_CHECK(ctx);
    } else {
      k_scan_blocksum<<<scan_blocks, SCAN_THREADS, 0, st>>>(counts, M, blocksums);
      VIMZ_LAUNCH_CHECK(ctx);
      k_scan_top<<<1, 1024, 0, st>>>(blocksums, scan_blocks, blocksums + scan_blocks + 1);
      VIMZ_LAUNCH_CHECK(ctx);
      k_scan_apply<<<scan_blocks, SCAN_THREADS, 0, st>>>(counts, M, blocksums, offsets, cursor);
      VIMZ_LAUNCH_CHECK(ctx);
    }
    if (n > 0) {
      k_msm_scatter<<<dim3(grid_n, nwin), 256, 0, st>>>(ws.digits.as<uint32_t>(), (uint32_t)n, nwin, (uint32_t)ck->n, (uint32_t)first, cursor, sorted,
                                            offsets, M, nthreads, seg_min, cb);
      VIMZ_LAUNCH_CHECK(ctx);
    }
  }
  if (ctx->prof.on && phase != 2) {  // total bucket insertions of this MSM = offsets[M]
    uint32_t* slot;
    if (!ctx->prof.entry_pool.empty()) { slot = ctx->prof.entry_pool.back(); ctx->prof.entry_pool.pop_back(); }
    else VIMZ_CUDA(cudaMallocHost(&slot, 4));
    VIMZ_CUDA(cudaMemcpyAsync(slot, offsets + M, 4, cudaMemcpyDeviceToHost, st));
    ctx->prof.entry_slots.push_back(slot);
    ctx->prof.entry_fused.push_back(counted);
  }
  {
    ProfScope prof_acc(ctx, PROF_MSM_ACCUMULATE, st);
    if (phase != 2) {
      if (ws.acc_wait) VIMZ_CUDA(cudaStreamWaitEvent(st, ws.acc_wait, 0));
      {
        ProfScope prof_kernel(ctx, PROF_MSM_ACC_KERNEL, st);
        ProfScope prof_fused(counted ? ctx : nullptr, PROF_MSM_ACC_KERNEL_FUSED, st);
        k_msm_accumulate<C><<<nthreads / 128, 128, 0, st>>>(offsets, sorted, ck->table, M, nthreads, seg_min, ws.buckets.ptr, ws.partials.ptr);
      }
      VIMZ_LAUNCH_CHECK(ctx);
      if (ws.acc_record) VIMZ_CUDA(cudaEventRecord(ws.acc_record, st));
    }
    if (phase == 1) return VIMZ_OK;
    // pieces of cut buckets: giants (blocks per chunk + last-arrival fold), mids (a warp each), the rest (a quad each)
    // (with deferred giants the giant role moves into k_reduce_tail, off the chain accumulate -> combine -> reduce)
    const uint32_t nb_big = defer ? 0u : (uint32_t)ctx->sm_count * VIMZ_BIG_BLOCKS_PER_SM, nb_mid = (uint32_t)ctx->sm_count * 4,
                   nb_small = ceil_div((size_t)M * 4, 128);
    k_msm_combine_all<C><<<nb_big + nb_mid + nb_small, 128, 0, st>>>(offsets, M, nthreads, seg_min, nb_big, nb_mid, ws.partials.ptr,
                                                                     ws.buckets.ptr, cb, defer);
    VIMZ_LAUNCH_CHECK(ctx);
  }
  ProfScope prof_red(ctx, PROF_MSM_REDUCE, st);
  k_reduce_chunks<C><<<ceil_div((size_t)T * 4, 128), 128, 0, st>>>(ws.buckets.ptr, T, K, ws.chunkA.ptr, ws.chunkL.ptr);
  VIMZ_LAUNCH_CHECK(ctx);
  // (the subtracted point is produced by another stream long before: inside a captured graph an EXTERNAL event wait, so that it
  // stays a wait on the event object and does not pull that stream into the capture)
  if (ws.sub_jac && ws.sub_event) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    VIMZ_CUDA(cudaStreamIsCapturing(st, &cs));
    VIMZ_CUDA(cudaStreamWaitEvent(st, ws.sub_event, cs == cudaStreamCaptureStatusActive ? cudaEventWaitExternal : cudaEventWaitDefault));
  }
  k_reduce_tail<C><<<dim3(G, nb + 1 + (defer ? 1 : 0)), 128, 0, st>>>(ws.chunkA.ptr, ws.chunkL.ptr, T, nb, logK, ws.bitsums.ptr, ws.scaled.ptr,
                                                                      cb.ctrl + CTRL_REDUCE, d_out, ws.host_out, ws.sub_jac, offsets, M, nthreads, seg_min, ws.partials.ptr, cb,
                                                                      defer ? ws.deferred.ptr : nullptr);
  VIMZ_LAUNCH_CHECK(ctx);
  return VIMZ_OK;
}

template <class C>
int impl_point_sum(vimz_ctx* ctx, const void* d_pts, size_t k, void* d_out) {
  k_point_sum<C><<<1, 32, 0, ctx->stream>>>(d_pts, (uint32_t)k, d_out);
  VIMZ_LAUNCH_CHECK(ctx);
  return VIMZ_OK;
}
template <class C>
int impl_point_add2(vimz_ctx* ctx, cudaStream_t st, const void* d_a, const void* d_b, void* d_out, void* host_out) {
  k_point_add2<C><<<1, 32, 0, st>>>(d_a, d_b, d_out, host_out);
  VIMZ_LAUNCH_CHECK(ctx);
  return VIMZ_OK;
}
template <class C>
int impl_point_sum_batch(vimz_ctx* ctx, const void* d_pts, size_t k, size_t sets, void* d_out) {
  if (sets == 0) return VIMZ_OK;
  k_point_sum_batch<C><<<(uint32_t)sets, 32, 0, ctx->stream>>>(d_pts, (uint32_t)k, (uint32_t)sets, d_out);
  VIMZ_LAUNCH_CHECK(ctx);
  return VIMZ_OK;
}
template <class C>
int impl_point_to_affine(vimz_ctx* ctx, const void* d_pt, void* d_out) {
  k_point_to_affine<C><<<1, 32, 0, ctx->stream>>>(d_pt, d_out);
  VIMZ_LAUNCH_CHECK(ctx);
  return VIMZ_OK;
}
template <class C>
int impl_point_scale_add(vimz_ctx* ctx, cudaStream_t st, const void* d_a, const void* d_r, const void* d_b, void* d_out, int count) {
  if (count > 8) return set_error(VIMZ_ERR_ARG, "point_scale_add: at most 8 pairs per call");
  k_point_scale_add<C><<<1, 32, 0, st>>>(d_a, d_r, d_b, d_out, count);
  VIMZ_LAUNCH_CHECK(ctx);
  return VIMZ_OK;
}
template <class C>
int impl_point_scale_add_val(vimz_ctx* ctx, cudaStream_t st, const void* d_a, const vimz_fr* r, const void* d_b, void* d_out, int count) {
  if (count > 8) return set_error(VIMZ_ERR_ARG, "point_scale_add: at most 8 pairs per call");
  Fp<typename C::Fs> rr;
  memcpy(rr.v, r, 32);
  k_point_scale_add_val<C><<<1, 32, 0, st>>>(d_a, rr, d_b, d_out, count);
  VIMZ_LAUNCH_CHECK(ctx);
  return VIMZ_OK;
}
template <class C>
int impl_gen_bases(vimz_ctx* ctx, uint64_t k0, uint64_t dk, size_t n, void* d_out) {
  if (n == 0) return VIMZ_OK;
  k_gen_bases<C><<<ceil_div(ceil_div(n, GEN_RUN), 128), 128, 0, ctx->stream>>>(k0, dk, (uint32_t)n, d_out);
  VIMZ_LAUNCH_CHECK(ctx);
  return VIMZ_OK;
}

inline CsrView csr_view(const vimz_shape* s, int k) {
  CsrView v;
  v.rowptr = s->rowptr[k];
  v.col = s->col[k];
  v.val = s->val[k];
  return v;
}

template <class C>
int impl_spmv3(vimz_ctx* ctx, const vimz_shape* s, const void* d_W, const void* d_tail, void* d_Az, void* d_Bz, void* d_Cz) {
  if (s->m == 0) return VIMZ_OK;
  ProfScope prof(ctx, PROF_SPMV, ctx->stream);
  k_spmv3<typename C::Fs><<<dim3(ceil_div(s->m, 256), 3), 256, 0, ctx->stream>>>(
      csr_view(s, 0), csr_view(s, 1), csr_view(s, 2), (uint32_t)s->m, (uint32_t)s->n, d_W, d_tail, d_Az, d_Bz, d_Cz);
  VIMZ_LAUNCH_CHECK(ctx);
  return VIMZ_OK;
}
template <class C>
int impl_cross_term(vimz_ctx* ctx, const vimz_shape* s, const void* d_W1, const void* d_tail1, const void* d_W2, const void* d_tail2, void* d_T,
                    const vimz_ck* fuse_ck, const void* cache1, void* cache2, const uint8_t* rowflag) {
  if (s->m == 0) return VIMZ_OK;
  DigitCount dc{nullptr, 0, 0, nullptr, 0};
  if (fuse_ck) {  // zero lane 0's histogram, then let the cross-term kernels fill it and the digit array
    const uint32_t M = 1u << (fuse_ck->c - 1);
    VIMZ_TRY(ctx->ws.digits.reserve(std::max<size_t>(s->m * (size_t)fuse_ck->nwin, 1) * 4));
    dc.digits = ctx->ws.digits.as<uint32_t>();
    dc.stride = s->m;
    if (!fuse_ck->dtable) {  // a direct-table key has no buckets: digits only
      VIMZ_TRY(msm_zero_control(ctx, ctx->ws, M, ctx->stream));
      dc.counts = ctx->ws.counts.as<uint32_t>() + msm_ctrl_words(ctx);
    }
    dc.c = fuse_ck->c;
    dc.nwin = fuse_ck->nwin;
  }
  ProfScope prof(ctx, PROF_CROSS_TERM, ctx->stream);
  CrossArgs ca;
  ca.A = csr_view(s, 0); ca.B = csr_view(s, 1); ca.Cm = csr_view(s, 2);
  ca.m = (uint32_t)s->m; ca.n = (uint32_t)s->n;
  ca.W1 = d_W1; ca.tail1 = d_tail1; ca.W2 = d_W2; ca.tail2 = d_tail2; ca.T = d_T; ca.dc = dc;
  if (ctx->opt_cross_stream && s->n_chunks) {
    MatvecStreamArgs ma;
    ma.A = ca.A; ma.B = ca.B; ma.Cm = ca.Cm;
    ma.m = ca.m; ma.n = ca.n;
    ma.stream = reinterpret_cast<const uint2*>(s->chunk_stream);
    ma.desc = reinterpret_cast<const ChunkDesc*>(s->chunk_desc);
    ma.dict = s->dict;
    const void* p1 = cache1;
    void* p2 = cache2;
    if (!(cache1 && cache2)) {  // stand-alone commit_T: both product triples are computed here, into context scratch
      const size_t mb3 = (size_t)3 * s->m * 32;
      VIMZ_TRY(ctx->tmp4.reserve(mb3));
      VIMZ_TRY(ctx->tmp5.reserve(mb3));
      ma.W = d_W1; ma.tail = d_tail1; ma.out = ctx->tmp4.ptr;
      k_matvec_stream<typename C::Fs><<<(uint32_t)s->n_chunks, 256, MATVEC_SMEM, ctx->stream>>>(ma);
      VIMZ_LAUNCH_CHECK(ctx);
      p1 = ctx->tmp4.ptr;
      p2 = ctx->tmp5.ptr;
    }
    ma.W = d_W2; ma.tail = d_tail2; ma.out = p2;
    k_matvec_stream<typename C::Fs><<<(uint32_t)s->n_chunks, 256, MATVEC_SMEM, ctx->stream>>>(ma);
    VIMZ_LAUNCH_CHECK(ctx);
    k_cross_finish<typename C::Fs><<<ceil_div(s->m, CROSS_FINISH_THREADS), CROSS_FINISH_THREADS, 0, ctx->stream>>>(p1, p2, d_tail1, (uint32_t)s->m, d_T, dc, rowflag);
    VIMZ_LAUNCH_CHECK(ctx);
    return VIMZ_OK;
  }
  if (rowflag) return set_error(VIMZ_ERR_ARG, "cross_term: the booleanity-row shift needs the streamed kernels (cross_stream = 1)");
  const uint32_t nb_long = ceil_div(s->n_long * 32, 128), nb_mid = ceil_div(s->n_mid * 8, 128), nb_short = ceil_div(s->m, 128);
  k_cross_term<typename C::Fs><<<nb_long + nb_mid + nb_short, 128, 0, ctx->stream>>>(ca, s->long_rows, (uint32_t)s->n_long, nb_long,
                                                                                     s->mid_rows, (uint32_t)s->n_mid, nb_mid);
  VIMZ_LAUNCH_CHECK(ctx);
  if (cache2) {  // the row-class kernel recomputes the z1 products itself; the accumulator still needs (Az2, Bz2, Cz2) for its fold
    char* c2 = reinterpret_cast<char*>(cache2);
    k_spmv3<typename C::Fs><<<dim3(ceil_div(s->m, 256), 3), 256, 0, ctx->stream>>>(
        csr_view(s, 0), csr_view(s, 1), csr_view(s, 2), (uint32_t)s->m, (uint32_t)s->n, d_W2, d_tail2, c2, c2 + s->m * 32, c2 + 2 * s->m * 32);
    VIMZ_LAUNCH_CHECK(ctx);
  }
  return VIMZ_OK;
}
template <class C>
int impl_mask_rows(vimz_ctx* ctx, cudaStream_t st, const void* d_v, const uint8_t* rowflag, size_t m, void* d_out) {
  if (m == 0) return VIMZ_OK;
  k_mask_rows<typename C::Fs><<<ceil_div(m, 256), 256, 0, st>>>(d_v, rowflag, (uint32_t)m, d_out);
  VIMZ_LAUNCH_CHECK(ctx);
  return VIMZ_OK;
}
// One 128-thread block on every second SM: ~55 k additions are ~6 per thread (~25 us) plus the tree, and a grid this thin leaves the
// kernels of the critical lane, which run beside it, their SMs (two blocks on every SM made k_cross_finish and k_msm_scatter 2x slower).
inline uint32_t masked_sum_blocks(const vimz_ctx* ctx) { return std::min<uint32_t>(std::max<uint32_t>((uint32_t)ctx->sm_count / 2, 1), DIRECT_MAX_BLOCKS); }
inline size_t masked_sum_scratch_bytes(const vimz_ctx* ctx) {
  const uint32_t blocks = masked_sum_blocks(ctx);
  return ((size_t)blocks + ceil_div(blocks, DIRECT_GROUP)) * 128 + 64 * 4;  // block sums, group sums, then the counters
}
template <class C>
int impl_masked_base_sum(vimz_ctx* ctx, cudaStream_t st, const void* d_vals, const uint32_t* bitcol, size_t m, const vimz_ck* ck, void* scratch,
                         void* d_out) {
  const uint32_t blocks = masked_sum_blocks(ctx);
  char* sc = reinterpret_cast<char*>(scratch);
  uint32_t* ctrl = reinterpret_cast<uint32_t*>(sc + masked_sum_scratch_bytes(ctx) - 64 * 4);
  k_masked_base_sum<C><<<blocks, 128, 0, st>>>(d_vals, bitcol, (uint32_t)m, ck->table, sc, ctrl, d_out);
  VIMZ_LAUNCH_CHECK(ctx);
  return VIMZ_OK;
}
template <class C>
int impl_point_pow2_parts(vimz_ctx* ctx, cudaStream_t st, const void* d_pt, void* d_parts) {
  k_point_pow2_parts<C><<<1, 32, 0, st>>>(d_pt, d_parts);
  VIMZ_LAUNCH_CHECK(ctx);
  return VIMZ_OK;
}
template <class C>
int impl_point_scale_add_parts(vimz_ctx* ctx, cudaStream_t st, const void* d_a, const vimz_fr* r, const void* d_parts, void* d_out) {
  Fp<typename C::Fs> rr;
  memcpy(rr.v, r, 32);
  k_point_scale_add_parts<C><<<1, 32 * SCALE_PARTS, 0, st>>>(d_a, rr, d_parts, d_out);
  VIMZ_LAUNCH_CHECK(ctx);
  return VIMZ_OK;
}
template <class C>
int impl_axpy(vimz_ctx* ctx, const void* d_a, const void* d_b, const vimz_fr* r, size_t len, void* d_out) {
  if (len == 0) return VIMZ_OK;
  Fp<typename C::Fs> rr;
  memcpy(rr.v, r, 32);
  int grid = (int)std::min<size_t>(ceil_div(len, 256), (size_t)ctx->sm_count * 16);
  ProfScope prof(ctx, PROF_AXPY, ctx->stream);
  k_axpy<typename C::Fs><<<grid, 256, 0, ctx->stream>>>(d_a, d_b, rr, len, d_out);
  VIMZ_LAUNCH_CHECK(ctx);
  return VIMZ_OK;
}
template <class C>
int impl_axpyn(vimz_ctx* ctx, const AxpySeg* segs, int count, const vimz_fr* r) {
  if (count < 1 || count > AXPY_MAX_SEGS) return set_error(VIMZ_ERR_ARG, "axpyn: 1..6 segments");
  AxpySegs sg;
  size_t total = 0;
  for (int k = 0; k < AXPY_MAX_SEGS; k++) {
    sg.s[k] = k < count ? segs[k] : AxpySeg{nullptr, nullptr, 0};
    total += sg.s[k].len;
    sg.end[k] = total;
  }
  sg.count = count;
  if (total == 0) return VIMZ_OK;
  Fp<typename C::Fs> rr;
  memcpy(rr.v, r, 32);
  int grid = (int)std::min<size_t>(ceil_div(total, 256), (size_t)ctx->sm_count * 16);
  ProfScope prof(ctx, PROF_AXPY, ctx->stream);
  k_axpy3<typename C::Fs><<<grid, 256, 0, ctx->stream>>>(sg, rr);
  VIMZ_LAUNCH_CHECK(ctx);
  return VIMZ_OK;
}
template <class C>
int impl_field_op(vimz_ctx* ctx, int which, int op, const void* d_a, const void* d_b, size_t n, void* d_out) {
  if (n == 0) return VIMZ_OK;
  if (which == 0)
    k_field_op<typename C::Fb><<<ceil_div(n, 256), 256, 0, ctx->stream>>>(op, d_a, d_b, n, d_out);
  else
    k_field_op<typename C::Fs><<<ceil_div(n, 256), 256, 0, ctx->stream>>>(op, d_a, d_b, n, d_out);
  VIMZ_LAUNCH_CHECK(ctx);
  return VIMZ_OK;
}

template <class C>
CurveVTable make_vtable(const char* name) {
  CurveVTable t;
  t.id = C::ID;
  t.name = name;
  for (int i = 0; i < 8; i++) {
    t.scalar_modulus[i] = C::Fs::p(i);
    t.scalar_one_mont[i] = C::Fs::one(i);
  }
  t.init_device = &impl_init_device<C>;
  t.precompute = &impl_precompute<C>;
  t.precompute_direct = &impl_precompute_direct<C>;
  t.msm = &impl_msm<C>;
  t.point_sum = &impl_point_sum<C>;
  t.point_sum_batch = &impl_point_sum_batch<C>;
  t.point_add2 = &impl_point_add2<C>;
  t.msm_digits = &impl_msm_digits<C>;
  t.point_to_affine = &impl_point_to_affine<C>;
  t.point_scale_add = &impl_point_scale_add<C>;
  t.point_scale_add_val = &impl_point_scale_add_val<C>;
  t.gen_bases = &impl_gen_bases<C>;
  t.spmv3 = &impl_spmv3<C>;
  t.cross_term = &impl_cross_term<C>;
  t.mask_rows = &impl_mask_rows<C>;
  t.masked_base_sum = &impl_masked_base_sum<C>;
  t.point_pow2_parts = &impl_point_pow2_parts<C>;
  t.point_scale_add_parts = &impl_point_scale_add_parts<C>;
  t.axpy = &impl_axpy<C>;
  t.axpyn = &impl_axpyn<C>;
  t.field_op = &impl_field_op<C>;
  return t;
}

}  // namespace vimz
