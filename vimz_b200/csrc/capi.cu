// capi.cu -- extern "C" entry points of libvimz_gpu.so (declared in include/vimz_gpu.h).
// Host logic only: argument checks with nova-snark's error behaviour, COO->CSR, window selection,
// staging copies and the launch sequences of curve_impl.cuh.  There is no CPU compute path: without
// a CUDA device vimz_ctx_create fails with VIMZ_ERR_NO_DEVICE.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <unordered_map>
#include <vector>
#include "curve_impl.cuh"

namespace vimz {
thread_local std::string g_last_error;
const CurveVTable* vtable_pallas();
const CurveVTable* vtable_vesta();
const CurveVTable* vtable_bn254();
const CurveVTable* vtable_grumpkin();

const CurveVTable* curve_vtable(int id) {
  switch (id) {
    case VIMZ_PALLAS: return vtable_pallas();
    case VIMZ_VESTA: return vtable_vesta();
    case VIMZ_BN254: return vtable_bn254();
    case VIMZ_GRUMPKIN: return vtable_grumpkin();
    default: return nullptr;
  }
}

// ---- window selection ---------------------------------------------------------------------------
static int bit_length(const uint32_t v[8]) {
  for (int i = 7; i >= 0; i--)
    if (v[i]) return 32 * i + 32 - __builtin_clz(v[i]);
  return 0;
}
// bits [pos, pos+len) of v, len <= 32
static uint64_t get_bits(const uint32_t v[8], int pos, int len) {
  uint64_t r = 0;
  for (int b = 0; b < len; b++) {
    int p = pos + b;
    if (p < 256 && ((v[p >> 5] >> (p & 31)) & 1)) r |= 1ull << b;
  }
  return r;
}
// true iff all bits of v in [0, pos) that are >= bit `from` are zero ... helper: v mod 2^pos < 2^from
static bool low_part_below(const uint32_t v[8], int pos, int from) {
  for (int p = from; p < pos; p++)
    if (p < 256 && ((v[p >> 5] >> (p & 31)) & 1)) return false;
  return true;
}
// Smallest window count for c-bit signed digits such that the top digit never exceeds 2^(c-1)
// for any scalar in [0, q).
int msm_num_windows(const uint32_t q[8], int c) {
  uint32_t qm1[8];
  memcpy(qm1, q, 32);
  for (int i = 0; i < 8; i++) {  // q - 1
    if (qm1[i]-- != 0) break;
  }
  int bits = bit_length(qm1);
  int w = (bits + c - 1) / c;
  if (w < 1) w = 1;
  uint64_t M = 1ull << (c - 1);
  uint64_t top = get_bits(qm1, c * (w - 1), c);
  bool ok = top < M;
  if (!ok && top == M && w >= 2) {
    // top digit can be M only if the lower part is < 2^(c(w-1)-1) for every such scalar: then
    // window w-2 is < M and produces no carry.
    ok = low_part_below(qm1, c * (w - 1), c * (w - 1) - 1);
  }
  return ok ? w : w + 1;
}

int msm_pick_window(const uint32_t q[8], size_t n) {
  int best = 8;
  double best_cost = 1e300;
  for (int c = 8; c <= 24; c++) {
    int w = msm_num_windows(q, c);
    if (w > MSM_MAX_WINDOWS) continue;
    if ((double)n * w >= 2147483648.0) continue;  // table index must fit 31 bits
    // in field multiplications: 10 per bucket insertion; the bucket reduction is latency-bound and costs
    // the equivalent of ~100 per bucket on B200 (measured: ~1.4 ns per bucket at 70 G modmul/s)
    double cost = (double)n * w * 10.0 + 100.0 * (double)(1ull << (c - 1));
    if (cost < best_cost) {
      best_cost = cost;
      best = c;
    }
  }
  return best;
}
}  // namespace vimz

using namespace vimz;

#define CHECK_ARG(cond, msg) \
  do {                       \
    if (!(cond)) return set_error(VIMZ_ERR_ARG, msg); \
  } while (0)

// Makes the context's device current for the duration of an entry point and restores the caller's device afterwards
// (a single-process multi-GPU host -- torch, a Rust prover with its own CUDA code -- must not find its current device changed).
struct DeviceGuard {
  int prev = -1, dev = -1;
  explicit DeviceGuard(int d) : dev(d) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) cudaSetDevice(dev);
  }
  ~DeviceGuard() {
    if (prev >= 0 && prev != dev) cudaSetDevice(prev);
  }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};
// context lock (see vimz_ctx::mu) + current device, for the duration of an entry point
struct CtxGuard {
  std::unique_lock<std::recursive_mutex> lock;
  DeviceGuard dev;
  explicit CtxGuard(vimz_ctx* ctx) : lock(ctx->mu), dev(ctx->device) {}
};

// every stream of a context: main, commitment folds, aux MSM lane, P_S branch, K_S update, early commitments
static cudaError_t sync_all_streams(vimz_ctx* ctx) {
  cudaError_t first = cudaSuccess;
  for (cudaStream_t st : {ctx->stream, ctx->side, ctx->aux, ctx->ps, ctx->ks, ctx->early})
    if (st) {
      cudaError_t e = cudaStreamSynchronize(st);
      if (first == cudaSuccess) first = e;
    }
  return first;
}

// ---- handle lifetimes ----------------------------------------------------------------------------------
// Children (keys, shapes, accumulators) hold a reference on their context, accumulators also on their shape and keys.
// A *_destroy call on a parent that still has children only drops the owner's reference: the object is freed when the
// last child goes, so destruction order on the host side (Drop order in Rust, garbage collection in Python) is free.
static void ctx_free(vimz_ctx* ctx) {
  DeviceGuard dg(ctx->device);
  sync_all_streams(ctx);
  ctx->ws.release();
  ctx->ws_aux.release();
  ctx->ws_early.release();
  ctx->tmp0.release(); ctx->tmp1.release(); ctx->tmp2.release();
  ctx->tmp3.release(); ctx->tmp4.release(); ctx->tmp5.release();
  if (ctx->pinned) cudaFreeHost(ctx->pinned);
  for (ProfSpan& sp : ctx->prof.open) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
  for (cudaEvent_t e : ctx->prof.pool) cudaEventDestroy(e);
  for (uint32_t* slot : ctx->prof.entry_slots) cudaFreeHost(slot);
  for (uint32_t* slot : ctx->prof.entry_pool) cudaFreeHost(slot);
  cudaStreamDestroy(ctx->stream);
  cudaStreamDestroy(ctx->side);
  cudaStreamDestroy(ctx->aux);
  cudaStreamDestroy(ctx->ps);
  cudaStreamDestroy(ctx->ks);
  cudaStreamDestroy(ctx->early);
  delete ctx;
}
static void ctx_release(vimz_ctx* ctx) {
  if (ctx->refs.fetch_sub(1) == 1) ctx_free(ctx);
}
static void ck_release(vimz_ck* ck) {
  if (ck->refs.fetch_sub(1) != 1) return;
  vimz_ctx* ctx = ck->ctx;
  {
    CtxGuard g(ctx);
    sync_all_streams(ctx);
    if (ck->table) cudaFree(ck->table);
    if (ck->dtable) cudaFree(ck->dtable);
    delete ck;
  }
  ctx_release(ctx);
}
static void shape_release(vimz_shape* s) {
  if (s->refs.fetch_sub(1) != 1) return;
  vimz_ctx* ctx = s->ctx;
  {
    CtxGuard g(ctx);
    cudaStreamSynchronize(ctx->stream);
    for (int k = 0; k < 3; k++) {
      if (s->rowptr[k]) cudaFree(s->rowptr[k]);
      if (s->col[k]) cudaFree(s->col[k]);
      if (s->val[k]) cudaFree(s->val[k]);
    }
    if (s->dict) cudaFree(s->dict);
    if (s->chunk_stream) cudaFree(s->chunk_stream);
    if (s->chunk_desc) cudaFree(s->chunk_desc);
    if (s->long_rows) cudaFree(s->long_rows);
    if (s->mid_rows) cudaFree(s->mid_rows);
    if (s->rowflag) cudaFree(s->rowflag);
    if (s->bitcol) cudaFree(s->bitcol);
    delete s;
  }
  ctx_release(ctx);
}

extern "C" {

const char* vimz_last_error(void) { return g_last_error.c_str(); }
int vimz_version(void) { return 100; }
void* vimz_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (bytes == 0 || cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) {
    cudaGetLastError();
    set_error(VIMZ_ERR_CUDA, "vimz_host_alloc: cudaHostAlloc failed");
    return nullptr;
  }
  return p;
}
void vimz_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

int vimz_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int vimz_ctx_create(int curve_id, int device, vimz_ctx** out) {
  CHECK_ARG(out != nullptr, "vimz_ctx_create: out is null");
  *out = nullptr;
  if (!curve_vtable(curve_id)) return set_error(VIMZ_ERR_ARG, "vimz_ctx_create: unknown curve id");
  int ndev = vimz_device_count();
  if (ndev <= 0) return set_error(VIMZ_ERR_NO_DEVICE, "vimz_ctx_create: no CUDA device visible (this library has no CPU path)");
  if (device < 0 || device >= ndev) return set_error(VIMZ_ERR_ARG, "vimz_ctx_create: device index out of range");
  DeviceGuard dg(device);
  vimz_ctx* ctx = new vimz_ctx();
  ctx->curve = curve_id;
  ctx->device = device;
  auto init = [&]() -> int {
    // The main stream carries the critical lane of a fold step (cross term -> commit(T)); the aux lane (commit(W2)) and the
    // side stream (commitment folds) have slack, so they get the lowest priority: when both lanes have a grid pending, the
    // block scheduler serves the critical one first.  VIMZ_STREAM_PRIORITY=0 disables it (A/B).
    int prio_lo = 0, prio_hi = 0;
    VIMZ_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));  // lo = numerically greatest = least urgent
    const char* pe = getenv("VIMZ_STREAM_PRIORITY");
    if (pe && atoi(pe) == 0) prio_hi = prio_lo;
    VIMZ_CUDA(cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, prio_hi));
    VIMZ_CUDA(cudaStreamCreateWithPriority(&ctx->side, cudaStreamNonBlocking, prio_lo));
    VIMZ_CUDA(cudaStreamCreateWithPriority(&ctx->aux, cudaStreamNonBlocking, prio_lo));
    VIMZ_CUDA(cudaStreamCreateWithPriority(&ctx->ps, cudaStreamNonBlocking, prio_lo));
    VIMZ_CUDA(cudaStreamCreateWithPriority(&ctx->ks, cudaStreamNonBlocking, prio_hi));
    VIMZ_CUDA(cudaStreamCreateWithPriority(&ctx->early, cudaStreamNonBlocking, prio_lo));
    VIMZ_CUDA(cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device));
    VIMZ_CUDA(cudaMallocHost(&ctx->pinned, 4096));
    VIMZ_TRY(ctx->ws.result.reserve(4096));
    return curve_vtable(curve_id)->init_device(ctx);
  };
  int rc = init();
  if (rc != VIMZ_OK) {
    std::string msg = g_last_error;
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->side) cudaStreamDestroy(ctx->side);
    if (ctx->aux) cudaStreamDestroy(ctx->aux);
    if (ctx->ps) cudaStreamDestroy(ctx->ps);
    if (ctx->ks) cudaStreamDestroy(ctx->ks);
    if (ctx->early) cudaStreamDestroy(ctx->early);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    ctx->ws.release();
    delete ctx;
    cudaGetLastError();
    return set_error(rc, msg);
  }
  *out = ctx;
  return VIMZ_OK;
}

void vimz_ctx_destroy(vimz_ctx* ctx) {
  if (!ctx) return;
  {  // wait for a call still running on another thread, then for the streams
    CtxGuard g(ctx);
    sync_all_streams(ctx);
  }
  ctx_release(ctx);  // freed now, or when the last key / shape / accumulator created on it is destroyed
}

int vimz_ctx_sync(vimz_ctx* ctx) {
  CHECK_ARG(ctx, "vimz_ctx_sync: null ctx");
  CtxGuard g(ctx);
  VIMZ_CUDA(sync_all_streams(ctx));
  return VIMZ_OK;
}

int vimz_ctx_set_option(vimz_ctx* ctx, const char* key, long value) {
  CHECK_ARG(ctx && key, "vimz_ctx_set_option: null argument");
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  if (strcmp(key, "msm_window") == 0) {
    if (value != 0 && (value < 2 || value > 24)) return set_error(VIMZ_ERR_ARG, "msm_window must be 0 (auto) or in [2, 24]");
    ctx->opt_window = value;
    return VIMZ_OK;
  }
  if (strcmp(key, "profile") == 0) {
    ctx->prof.on = value != 0;
    return VIMZ_OK;
  }
  if (strcmp(key, "msm_acc_blocks") == 0) {
    if (value < 1 || value > 8) return set_error(VIMZ_ERR_ARG, "msm_acc_blocks must be in [1, 8]");
    ctx->opt_acc_blocks = value;
    alloc_epoch()++;  // captured step graphs hold the old launch geometry: rebuild them
    return VIMZ_OK;
  }
  if (strcmp(key, "msm_seg_min") == 0) {
    if (value < 1 || value > 4096) return set_error(VIMZ_ERR_ARG, "msm_seg_min must be in [1, 4096]");
    ctx->opt_seg_min = value;
    alloc_epoch()++;
    return VIMZ_OK;
  }
  if (strcmp(key, "msm_seg_min_aux") == 0) {  // the same for the MSMs of lane 1 (a fold step's commit(W2)); 0 = msm_seg_min
    if (value < 0 || value > 4096) return set_error(VIMZ_ERR_ARG, "msm_seg_min_aux must be in [0, 4096]");
    ctx->opt_seg_min_aux = value;
    alloc_epoch()++;
    return VIMZ_OK;
  }
  if (strcmp(key, "cross_stream") == 0) {
    ctx->opt_cross_stream = value != 0;
    alloc_epoch()++;
    return VIMZ_OK;
  }
  if (strcmp(key, "msm_direct_max") == 0) {  // applies to keys uploaded afterwards
    if (value < 0 || value > (1 << 20)) return set_error(VIMZ_ERR_ARG, "msm_direct_max must be in [0, 2^20]");
    ctx->opt_direct_max = value;
    return VIMZ_OK;
  }
  if (strcmp(key, "msm_direct_c") == 0) {  // applies to keys uploaded afterwards; 0 = by key length
    if (value != 0 && (value < 4 || value > 14)) return set_error(VIMZ_ERR_ARG, "msm_direct_c must be 0 or in [4, 14]");
    ctx->opt_direct_c = value;
    return VIMZ_OK;
  }
  if (strcmp(key, "msm_direct_bps") == 0) {
    if (value < 1 || value > 4) return set_error(VIMZ_ERR_ARG, "msm_direct_bps must be in [1, 4]");
    ctx->opt_direct_bps = value;
    alloc_epoch()++;
    return VIMZ_OK;
  }
  if (strcmp(key, "cross_cache") == 0) {  // applies to accumulators created afterwards
    ctx->opt_cross_cache = value != 0;
    return VIMZ_OK;
  }
  if (strcmp(key, "aux_lane") == 0) {
    ctx->opt_aux_lane = value != 0;
    alloc_epoch()++;  // changes the captured launch sequence
    return VIMZ_OK;
  }
  if (strcmp(key, "graph") == 0) {
    ctx->opt_graph = value != 0;
    return VIMZ_OK;
  }
  if (strcmp(key, "msm_defer_giants") == 0) {
    ctx->opt_defer_giants = value != 0;
    alloc_epoch()++;
    return VIMZ_OK;
  }
  if (strcmp(key, "acc_order") == 0) {
    ctx->opt_acc_order = value != 0;
    alloc_epoch()++;  // (shapes the captured launch sequence)
    return VIMZ_OK;
  }
  if (strcmp(key, "stage_commit") == 0) {
    ctx->opt_stage_commit = value != 0;
    return VIMZ_OK;
  }
  if (strcmp(key, "bitrow_fold") == 0) {  // applies to accumulators created afterwards
    ctx->opt_bitrow_fold = value != 0;
    return VIMZ_OK;
  }
  if (strcmp(key, "spin_wait") == 0) {
    ctx->opt_spin_wait = value != 0;
    return VIMZ_OK;
  }
  return set_error(VIMZ_ERR_ARG, std::string("unknown option: ") + key);
}

static const char* PROF_NAMES[PROF_COUNT] = {"msm_sort", "msm_accumulate", "msm_reduce", "cross_term", "axpy", "spmv", "msm_accumulate_kernel",
                                             "msm_accumulate_kernel_T"};

int vimz_ctx_profile(vimz_ctx* ctx, const char* name, double* ms, uint64_t* calls, int reset) {
  CHECK_ARG(ctx && name, "vimz_ctx_profile: null argument");
  CtxGuard g(ctx);
  VIMZ_CUDA(sync_all_streams(ctx));
  Profiler& p = ctx->prof;
  for (ProfSpan& s : p.open) {
    float t = 0;
    if (cudaEventElapsedTime(&t, s.a, s.b) == cudaSuccess) {
      p.ms[s.timer] += t;
      p.calls[s.timer]++;
    }
    p.pool.push_back(s.a);
    p.pool.push_back(s.b);
  }
  p.open.clear();
  for (size_t k = 0; k < p.entry_slots.size(); k++) {
    uint32_t* slot = p.entry_slots[k];
    p.msm_entries += *slot;
    if (k < p.entry_fused.size() && p.entry_fused[k]) p.msm_entries_fused += *slot;
    p.entry_pool.push_back(slot);
  }
  p.entry_slots.clear();
  p.entry_fused.clear();
  int rc = VIMZ_ERR_ARG;
  if (strcmp(name, "msm_entries") == 0) {
    if (ms) *ms = 0;
    if (calls) *calls = p.msm_entries;
    rc = VIMZ_OK;
  }
  if (strcmp(name, "msm_entries_T") == 0) {  // insertions of the commit(T) launches alone
    if (ms) *ms = 0;
    if (calls) *calls = p.msm_entries_fused;
    rc = VIMZ_OK;
  }
  // statistics of the last MSM of a lane: "lane0_entries" (bucket insertions), "lane0_nmid" / "lane0_ngiant" /
  // "lane0_nchunk" (buckets / chunks handled by the warp and block roles of k_msm_combine_all); same for lane1
  if (strncmp(name, "lane", 4) == 0 && (name[4] == '0' || name[4] == '1') && name[5] == '_') {
    MsmWorkspace& ws = name[4] == '0' ? ctx->ws : ctx->ws_aux;
    const char* what = name + 6;
    uint32_t v = 0;
    const uint32_t* src = nullptr;
    if (ws.last_M && ws.counts.ptr && ws.offsets.ptr) {  // the control block sits in front of the histogram
      if (strcmp(what, "entries") == 0) src = ws.offsets.as<uint32_t>() + ws.last_M;
      else if (strcmp(what, "nmid") == 0) src = ws.counts.as<uint32_t>() + CTRL_NMID;
      else if (strcmp(what, "ngiant") == 0) src = ws.counts.as<uint32_t>() + CTRL_NGIANT;
      else if (strcmp(what, "nchunk") == 0) src = ws.counts.as<uint32_t>() + CTRL_NCHUNK;
    }
    if (src) {
      VIMZ_CUDA(cudaMemcpy(&v, src, 4, cudaMemcpyDeviceToHost));
      if (ms) *ms = 0;
      if (calls) *calls = v;
      rc = VIMZ_OK;
    }
  }
  for (int k = 0; k < PROF_COUNT && rc != VIMZ_OK; k++)
    if (strcmp(name, PROF_NAMES[k]) == 0) {
      if (ms) *ms = p.ms[k];
      if (calls) *calls = p.calls[k];
      rc = VIMZ_OK;
    }
  if (reset) {
    for (int k = 0; k < PROF_COUNT; k++) { p.ms[k] = 0; p.calls[k] = 0; }
    p.msm_entries = 0;
    p.msm_entries_fused = 0;
  }
  return rc == VIMZ_OK ? rc : set_error(VIMZ_ERR_ARG, std::string("unknown profile timer: ") + name);
}

void* vimz_ctx_stream(vimz_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
uint64_t vimz_ctx_launch_count(vimz_ctx* ctx) { return ctx ? ctx->launches : 0; }

// ---- commitment key --------------------------------------------------------------------------------
static int ck_build(vimz_ctx* ctx, const void* d_bases, size_t n, vimz_ck** out) {
  const CurveVTable* vt = curve_vtable(ctx->curve);
  // short keys (the secondary curve of the fold, the hash / redact circuits) take the direct multiples table unless a
  // window size was forced; if its allocation fails the key falls back to the bucket pipeline
  bool direct = !ctx->opt_window && n > 0 && (long)n <= ctx->opt_direct_max;
  // 2^(c-1) multiples per (window, point): c = 10 (26 windows, 832 KB per point) up to 16 384 points, c = 8 (32 windows, 256 KB) above
  const int direct_c = ctx->opt_direct_c ? (int)ctx->opt_direct_c : (n <= 16384 ? DIRECT_C_SHORT : DIRECT_C);
  int c = ctx->opt_window ? (int)ctx->opt_window : (direct ? direct_c : msm_pick_window(vt->scalar_modulus, n));
  int nwin = msm_num_windows(vt->scalar_modulus, c);
  if (direct && nwin > MSM_MAX_WINDOWS) {
    direct = false;
    c = msm_pick_window(vt->scalar_modulus, n);
    nwin = msm_num_windows(vt->scalar_modulus, c);
  }
  void* dtable = nullptr;
  if (direct && cudaMalloc(&dtable, (n * (size_t)nwin * 64) << (c - 1)) != cudaSuccess) {
    cudaGetLastError();
    dtable = nullptr;
    c = msm_pick_window(vt->scalar_modulus, n);
    nwin = msm_num_windows(vt->scalar_modulus, c);
  }
  if (nwin > MSM_MAX_WINDOWS) return set_error(VIMZ_ERR_ARG, "msm_window too small: more than 32 windows");
  if ((double)n * nwin >= 2147483648.0) return set_error(VIMZ_ERR_ARG, "commitment key too long for this window size");
  vimz_ck* ck = new vimz_ck();
  ck->ctx = ctx;
  ctx->refs++;
  ck->n = n;
  ck->c = c;
  ck->nwin = nwin;
  size_t bytes = std::max<size_t>(n * (size_t)nwin * 64, 64);
  cudaError_t e = cudaMalloc(&ck->table, bytes);
  if (e != cudaSuccess) {
    if (dtable) cudaFree(dtable);
    delete ck;
    ctx->refs--;
    return set_error(VIMZ_ERR_CUDA, std::string("cudaMalloc(window table) failed: ") + cudaGetErrorString(e));
  }
  ck->dtable = dtable;
  int rc = vt->precompute(ctx, d_bases, n, c, nwin, ck->table);
  if (rc == VIMZ_OK && dtable) rc = vt->precompute_direct(ctx, ck->table, n, c, nwin, dtable);
  if (rc == VIMZ_OK) {
    cudaError_t s = cudaStreamSynchronize(ctx->stream);
    if (s != cudaSuccess) rc = set_error(VIMZ_ERR_CUDA, std::string("window-table expansion failed: ") + cudaGetErrorString(s));
  }
  if (rc != VIMZ_OK) {
    cudaFree(ck->table);
    if (ck->dtable) cudaFree(ck->dtable);
    delete ck;
    ctx->refs--;
    return rc;
  }
  *out = ck;
  return VIMZ_OK;
}

int vimz_ck_upload_dev(vimz_ctx* ctx, const void* d_bases, size_t n, vimz_ck** out) {
  CHECK_ARG(ctx && out && (d_bases || n == 0), "vimz_ck_upload_dev: null argument");
  CtxGuard g(ctx);
  return ck_build(ctx, d_bases, n, out);
}

int vimz_ck_upload(vimz_ctx* ctx, const vimz_affine* bases, size_t n, vimz_ck** out) {
  CHECK_ARG(ctx && out && (bases || n == 0), "vimz_ck_upload: null argument");
  CtxGuard g(ctx);
  void* d = nullptr;
  VIMZ_CUDA(cudaMalloc(&d, std::max<size_t>(n * 64, 64)));
  cudaError_t e = cudaMemcpyAsync(d, bases, n * 64, cudaMemcpyHostToDevice, ctx->stream);
  int rc = e == cudaSuccess ? ck_build(ctx, d, n, out) : set_error(VIMZ_ERR_CUDA, cudaGetErrorString(e));
  cudaStreamSynchronize(ctx->stream);
  cudaFree(d);
  return rc;
}

void vimz_ck_destroy(vimz_ck* ck) {
  if (ck) ck_release(ck);  // freed when no accumulator uses it any more
}
size_t vimz_ck_len(const vimz_ck* ck) { return ck ? ck->n : 0; }
int vimz_ck_window_bits(const vimz_ck* ck) { return ck ? ck->c : 0; }
int vimz_ck_num_windows(const vimz_ck* ck) { return ck ? ck->nwin : 0; }

// ---- MSM ---------------------------------------------------------------------------------------------
int vimz_msm_async_dev(vimz_ctx* ctx, const vimz_ck* ck, size_t first, const void* d_scalars, size_t n, void* d_out) {
  CHECK_ARG(ctx && ck && d_out && (d_scalars || n == 0), "vimz_msm: null argument");
  CHECK_ARG(ck->ctx == ctx, "vimz_msm: commitment key belongs to another context");
  if (first + n > ck->n) return set_error(VIMZ_ERR_LENGTH, "vimz_msm: vector longer than the commitment key");
  CtxGuard g(ctx);
  return curve_vtable(ctx->curve)->msm(ctx, 0, ck, first, d_scalars, n, d_out, false);
}

static int fetch_points(vimz_ctx* ctx, const void* d_src, void* host_dst, size_t bytes) {
  VIMZ_CUDA(cudaMemcpyAsync(ctx->pinned, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  VIMZ_CUDA(cudaStreamSynchronize(ctx->stream));
  memcpy(host_dst, ctx->pinned, bytes);
  return VIMZ_OK;
}

int vimz_msm_range_dev(vimz_ctx* ctx, const vimz_ck* ck, size_t first, const void* d_scalars, size_t n, vimz_point* out) {
  CHECK_ARG(ctx && out, "vimz_msm: null argument");
  // one lock around enqueue AND fetch: two host threads committing on the same context share ws.result, and with the
  // lock dropped in between, thread A could read thread B's commitment (the mutex is recursive, the nested call is fine)
  CtxGuard g(ctx);
  VIMZ_TRY(vimz_msm_async_dev(ctx, ck, first, d_scalars, n, ctx->ws.result.ptr));
  return fetch_points(ctx, ctx->ws.result.ptr, out, 96);
}

int vimz_msm_dev(vimz_ctx* ctx, const vimz_ck* ck, const void* d_scalars, size_t n, vimz_point* out) {
  return vimz_msm_range_dev(ctx, ck, 0, d_scalars, n, out);
}

int vimz_msm(vimz_ctx* ctx, const vimz_ck* ck, const vimz_fr* scalars, size_t n, vimz_point* out) {
  CHECK_ARG(ctx && ck && out && (scalars || n == 0), "vimz_msm: null argument");
  if (n > ck->n) return set_error(VIMZ_ERR_LENGTH, "vimz_msm: vector longer than the commitment key");
  CtxGuard g(ctx);
  VIMZ_TRY(ctx->ws.scal.reserve(std::max<size_t>(n * 32, 32)));
  VIMZ_CUDA(cudaMemcpyAsync(ctx->ws.scal.ptr, scalars, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  return vimz_msm_range_dev(ctx, ck, 0, ctx->ws.scal.ptr, n, out);
}

// ---- group helpers ---------------------------------------------------------------------------------
int vimz_point_sum(vimz_ctx* ctx, const vimz_point* pts, size_t k, vimz_point* out) {
  CHECK_ARG(ctx && out && (pts || k == 0), "vimz_point_sum: null argument");
  CtxGuard g(ctx);
  VIMZ_TRY(ctx->tmp0.reserve(std::max<size_t>(k * 96, 96) + 96));
  char* d = ctx->tmp0.as<char>();
  VIMZ_CUDA(cudaMemcpyAsync(d + 96, pts, k * 96, cudaMemcpyHostToDevice, ctx->stream));
  VIMZ_TRY(curve_vtable(ctx->curve)->point_sum(ctx, d + 96, k, d));
  return fetch_points(ctx, d, out, 96);
}

int vimz_point_to_affine(vimz_ctx* ctx, const vimz_point* p, vimz_affine* out) {
  CHECK_ARG(ctx && p && out, "vimz_point_to_affine: null argument");
  CtxGuard g(ctx);
  VIMZ_TRY(ctx->tmp0.reserve(256));
  char* d = ctx->tmp0.as<char>();
  VIMZ_CUDA(cudaMemcpyAsync(d, p, 96, cudaMemcpyHostToDevice, ctx->stream));
  VIMZ_TRY(curve_vtable(ctx->curve)->point_to_affine(ctx, d, d + 128));
  return fetch_points(ctx, d + 128, out, 64);
}

int vimz_point_scale_add(vimz_ctx* ctx, const vimz_point* a, const vimz_fr* r, const vimz_point* b, vimz_point* out) {
  CHECK_ARG(ctx && a && r && b && out, "vimz_point_scale_add: null argument");
  CtxGuard g(ctx);
  VIMZ_TRY(ctx->tmp0.reserve(512));
  char* d = ctx->tmp0.as<char>();
  VIMZ_CUDA(cudaMemcpyAsync(d, a, 96, cudaMemcpyHostToDevice, ctx->stream));
  VIMZ_CUDA(cudaMemcpyAsync(d + 96, b, 96, cudaMemcpyHostToDevice, ctx->stream));
  VIMZ_CUDA(cudaMemcpyAsync(d + 192, r, 32, cudaMemcpyHostToDevice, ctx->stream));
  VIMZ_TRY(curve_vtable(ctx->curve)->point_scale_add(ctx, ctx->stream, d, d + 192, d + 96, d + 256, 1));
  return fetch_points(ctx, d + 256, out, 96);
}

// ---- R1CS shape --------------------------------------------------------------------------------------
// Distinct coefficient values of a shape.  R1CS matrices repeat a handful of constants (1, -1, powers of two,
// Poseidon round / MDS constants), so the streamed cross term reads a 4-byte index per non-zero instead of 32 bytes.
struct FrHash {
  size_t operator()(const vimz_fr& v) const { return (size_t)(v.l[0] * 0x9E3779B97F4A7C15ull ^ v.l[1] * 0xC2B2AE3D27D4EB4Full ^ v.l[2] ^ (v.l[3] << 1)); }
};
struct FrEq {
  bool operator()(const vimz_fr& a, const vimz_fr& b) const { return memcmp(&a, &b, 32) == 0; }
};
struct ValueDict {
  std::unordered_map<vimz_fr, uint32_t, FrHash, FrEq> index;
  std::vector<vimz_fr> values;
  uint32_t lookup(const vimz_fr& v) {
    auto it = index.find(v);
    if (it != index.end()) return it->second;
    uint32_t id = (uint32_t)values.size();
    values.push_back(v);
    index.emplace(v, id);
    return id;
  }
};

static int coo_to_csr(vimz_ctx* ctx, size_t m, size_t ncols, const uint32_t* row, const uint32_t* col, const vimz_fr* val, size_t nnz,
                      uint32_t** d_rowptr, uint32_t** d_col, void** d_val, std::vector<uint32_t>& row_nnz, ValueDict& dict,
                      std::vector<uint32_t>& h_rowptr, std::vector<uint32_t>& h_col, std::vector<uint32_t>& h_vidx) {
  std::vector<uint32_t> rowptr(m + 1, 0);
  for (size_t k = 0; k < nnz; k++) {
    if (row[k] >= m || col[k] >= ncols) return set_error(VIMZ_ERR_INDEX, "vimz_shape_upload: entry out of range (InvalidIndex)");
    rowptr[row[k] + 1]++;
  }
  for (size_t i = 0; i < m; i++) row_nnz[i] += rowptr[i + 1];
  for (size_t i = 0; i < m; i++) rowptr[i + 1] += rowptr[i];
  std::vector<uint32_t> cursor(rowptr.begin(), rowptr.end() - 1), ccol(nnz);
  std::vector<vimz_fr> cval(nnz);
  for (size_t k = 0; k < nnz; k++) {  // stable: keeps constraint order inside a row
    uint32_t p = cursor[row[k]]++;
    ccol[p] = col[k];
    cval[p] = val[k];
  }
  h_vidx.resize(nnz);
  for (size_t k = 0; k < nnz; k++) h_vidx[k] = dict.lookup(cval[k]);
  VIMZ_CUDA(cudaMalloc(d_rowptr, (m + 1) * 4));
  VIMZ_CUDA(cudaMalloc(d_col, std::max<size_t>(nnz * 4, 4)));
  VIMZ_CUDA(cudaMalloc(d_val, std::max<size_t>(nnz * 32, 32)));
  VIMZ_CUDA(cudaMemcpyAsync(*d_rowptr, rowptr.data(), (m + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
  VIMZ_CUDA(cudaMemcpyAsync(*d_col, ccol.data(), nnz * 4, cudaMemcpyHostToDevice, ctx->stream));
  VIMZ_CUDA(cudaMemcpyAsync(*d_val, cval.data(), nnz * 32, cudaMemcpyHostToDevice, ctx->stream));
  VIMZ_CUDA(cudaStreamSynchronize(ctx->stream));
  h_rowptr.swap(rowptr);
  h_col.swap(ccol);
  return VIMZ_OK;
}

void vimz_shape_destroy(vimz_shape* s) {
  if (s) shape_release(s);  // freed when no accumulator uses it any more
}

int vimz_shape_upload(vimz_ctx* ctx, size_t num_cons, size_t num_vars, size_t num_io,
                      const uint32_t* rowA, const uint32_t* colA, const vimz_fr* valA, size_t nnzA,
                      const uint32_t* rowB, const uint32_t* colB, const vimz_fr* valB, size_t nnzB,
                      const uint32_t* rowC, const uint32_t* colC, const vimz_fr* valC, size_t nnzC,
                      vimz_shape** out) {
  CHECK_ARG(ctx && out, "vimz_shape_upload: null argument");
  CHECK_ARG((nnzA == 0 || (rowA && colA && valA)) && (nnzB == 0 || (rowB && colB && valB)) && (nnzC == 0 || (rowC && colC && valC)),
            "vimz_shape_upload: null matrix arrays");
  CHECK_ARG(num_cons < (1ull << 31) && num_vars + num_io + 1 < (1ull << 31), "vimz_shape_upload: shape too large");
  CtxGuard g(ctx);
  vimz_shape* s = new vimz_shape();
  s->ctx = ctx;
  ctx->refs++;
  s->m = num_cons;
  s->n = num_vars;
  s->io = num_io;
  size_t ncols = num_vars + 1 + num_io;
  const uint32_t* rows[3] = {rowA, rowB, rowC};
  const uint32_t* cols[3] = {colA, colB, colC};
  const vimz_fr* vals[3] = {valA, valB, valC};
  size_t nnz[3] = {nnzA, nnzB, nnzC};
  std::vector<uint32_t> row_nnz(num_cons, 0);
  ValueDict dict;
  {  // fixed slots: 0 = +1, 1 = -1 (Montgomery form) -- the kernels short-cut both
    const CurveVTable* vt = curve_vtable(ctx->curve);
    vimz_fr one, minus_one;
    memcpy(&one, vt->scalar_one_mont, 32);
    unsigned __int128 borrow = 0;
    const uint64_t* q64 = reinterpret_cast<const uint64_t*>(vt->scalar_modulus);
    for (int i = 0; i < 4; i++) {  // q - one
      unsigned __int128 d = (unsigned __int128)q64[i] - one.l[i] - (uint64_t)borrow;
      minus_one.l[i] = (uint64_t)d;
      borrow = (d >> 64) & 1;
    }
    dict.lookup(one);
    dict.lookup(minus_one);
  }
  std::vector<uint32_t> h_rowptr[3], h_col[3], h_vidx[3];
  for (int k = 0; k < 3; k++) {
    s->nnz[k] = nnz[k];
    int rc = coo_to_csr(ctx, num_cons, ncols, rows[k], cols[k], vals[k], nnz[k], &s->rowptr[k], &s->col[k], &s->val[k], row_nnz, dict,
                        h_rowptr[k], h_col[k], h_vidx[k]);
    if (rc != VIMZ_OK) {
      vimz_shape_destroy(s);
      return rc;
    }
  }
  // rows too long for one thread get 8 lanes or a whole warp in the cross-term kernels
  std::vector<uint32_t> long_rows, mid_rows;
  for (size_t i = 0; i < num_cons; i++) {
    if (row_nnz[i] > R1CS_LONG_ROW) long_rows.push_back((uint32_t)i);
    else if (row_nnz[i] > R1CS_SHORT_ROW) mid_rows.push_back((uint32_t)i);
  }
  s->n_long = long_rows.size();
  s->n_mid = mid_rows.size();
  cudaError_t e = cudaSuccess;
  if (s->n_long) {
    e = cudaMalloc(&s->long_rows, s->n_long * 4);
    if (e == cudaSuccess) e = cudaMemcpy(s->long_rows, long_rows.data(), s->n_long * 4, cudaMemcpyHostToDevice);
  }
  if (e == cudaSuccess && s->n_mid) {
    e = cudaMalloc(&s->mid_rows, s->n_mid * 4);
    if (e == cudaSuccess) e = cudaMemcpy(s->mid_rows, mid_rows.data(), s->n_mid * 4, cudaMemcpyHostToDevice);
  }
  // streamed mat-vec: consecutive rows are grouped into chunks of <= CROSS_CHUNK_NNZ non-zeros (A+B+C) and
  // <= CROSS_CHUNK_ROWS rows; a row above CROSS_ROW_MAX non-zeros is a chunk of its own, walked by one warp in global memory.
  // Every other chunk gets its (col, vidx) pairs packed contiguously -- A's entries, then B's, then C's, each in CSR order --
  // starting at an even pair index (16-byte aligned: the kernel fetches the stream with one TMA bulk copy).
  std::vector<ChunkDesc> desc;
  std::vector<uint2> stream;
  {
    size_t i = 0;
    while (i < num_cons) {
      ChunkDesc d{};
      if (row_nnz[i] > CROSS_ROW_MAX) {
        d.off = CHUNK_LONG_ROW; d.r0 = (uint32_t)i; d.nrows = 1;
        desc.push_back(d);
        i++;
        continue;
      }
      size_t acc = 0, j = i;
      while (j < num_cons && j - i < CROSS_CHUNK_ROWS && row_nnz[j] <= CROSS_ROW_MAX && acc + row_nnz[j] <= CROSS_CHUNK_NNZ) acc += row_nnz[j++];
      if (stream.size() & 1) stream.push_back(make_uint2(0, 0));
      d.off = (uint32_t)stream.size(); d.r0 = (uint32_t)i; d.nrows = (uint32_t)(j - i);
      uint32_t* cnt[3] = {&d.nA, &d.nB, &d.nC};
      for (int k = 0; k < 3; k++) {
        const uint32_t b = h_rowptr[k][i], e2 = h_rowptr[k][j];
        *cnt[k] = e2 - b;
        for (uint32_t t = b; t < e2; t++) stream.push_back(make_uint2(h_col[k][t], h_vidx[k][t]));
      }
      desc.push_back(d);
      i = j;
    }
    s->n_chunks = desc.size();
    stream.push_back(make_uint2(0, 0));  // the bulk copy rounds a chunk's byte count up to 16
    stream.push_back(make_uint2(0, 0));
    if (stream.size() >= (1ull << 32)) {
      vimz_shape_destroy(s);
      return set_error(VIMZ_ERR_ARG, "vimz_shape_upload: shape too large (packed index stream)");
    }
  }
  // booleanity rows b * (b - 1) = 0: A = {(b, +1)}, B = {(b, +1), (one, -1)} in either order, C = {} -- `one` is column num_vars
  // (the u slot of z = (W, u, X)); dictionary slots 0 / 1 are +1 / -1.  See k_cross_finish for what the accumulator does with them.
  std::vector<uint8_t> rowflag(num_cons, 0);
  std::vector<uint32_t> bitcol(num_cons, 0xffffffffu);
  for (size_t i = 0; i < num_cons; i++) {
    const uint32_t a0 = h_rowptr[0][i], a1 = h_rowptr[0][i + 1], b0 = h_rowptr[1][i], b1 = h_rowptr[1][i + 1];
    if (a1 - a0 != 1 || b1 - b0 != 2 || h_rowptr[2][i + 1] != h_rowptr[2][i]) continue;
    const uint32_t bcol = h_col[0][a0];
    if (bcol >= num_vars || h_vidx[0][a0] != 0) continue;
    bool ok = false;
    for (int k = 0; k < 2 && !ok; k++) {
      const uint32_t p = b0 + k, q2 = b0 + 1 - k;
      ok = h_col[1][p] == bcol && h_vidx[1][p] == 0 && h_col[1][q2] == (uint32_t)num_vars && h_vidx[1][q2] == 1;
    }
    if (ok) { rowflag[i] = 1; bitcol[i] = bcol; s->n_bitrows++; }
  }
  s->n_dict = dict.values.size();
  if (e == cudaSuccess) e = cudaMalloc(&s->chunk_stream, stream.size() * sizeof(uint2));
  if (e == cudaSuccess) e = cudaMemcpy(s->chunk_stream, stream.data(), stream.size() * sizeof(uint2), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMalloc(&s->chunk_desc, std::max<size_t>(desc.size(), 1) * sizeof(ChunkDesc));
  if (e == cudaSuccess && !desc.empty()) e = cudaMemcpy(s->chunk_desc, desc.data(), desc.size() * sizeof(ChunkDesc), cudaMemcpyHostToDevice);
  if (e == cudaSuccess && num_cons) e = cudaMalloc(&s->rowflag, num_cons);
  if (e == cudaSuccess && num_cons) e = cudaMemcpy(s->rowflag, rowflag.data(), num_cons, cudaMemcpyHostToDevice);
  if (e == cudaSuccess && num_cons) e = cudaMalloc(&s->bitcol, num_cons * sizeof(uint32_t));
  if (e == cudaSuccess && num_cons) e = cudaMemcpy(s->bitcol, bitcol.data(), num_cons * sizeof(uint32_t), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMalloc(&s->dict, s->n_dict * 32);
  if (e == cudaSuccess) e = cudaMemcpy(s->dict, dict.values.data(), s->n_dict * 32, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    vimz_shape_destroy(s);
    return set_error(VIMZ_ERR_CUDA, std::string("vimz_shape_upload: ") + cudaGetErrorString(e));
  }
  *out = s;
  return VIMZ_OK;
}

int vimz_multiply_vec(vimz_ctx* ctx, const vimz_shape* s, const vimz_fr* z, size_t z_len, vimz_fr* Az, vimz_fr* Bz, vimz_fr* Cz) {
  CHECK_ARG(ctx && s && z && Az && Bz && Cz, "vimz_multiply_vec: null argument");
  if (z_len != s->n + 1 + s->io) return set_error(VIMZ_ERR_LENGTH, "vimz_multiply_vec: z.len() != num_io + num_vars + 1 (InvalidWitnessLength)");
  CtxGuard g(ctx);
  size_t mb = std::max<size_t>(s->m * 32, 32);
  VIMZ_TRY(ctx->tmp0.reserve(z_len * 32));
  VIMZ_TRY(ctx->tmp1.reserve(mb));
  VIMZ_TRY(ctx->tmp2.reserve(mb));
  VIMZ_TRY(ctx->tmp3.reserve(mb));
  char* dz = ctx->tmp0.as<char>();
  VIMZ_CUDA(cudaMemcpyAsync(dz, z, z_len * 32, cudaMemcpyHostToDevice, ctx->stream));
  VIMZ_TRY(curve_vtable(ctx->curve)->spmv3(ctx, s, dz, dz + s->n * 32, ctx->tmp1.ptr, ctx->tmp2.ptr, ctx->tmp3.ptr));
  VIMZ_CUDA(cudaMemcpyAsync(Az, ctx->tmp1.ptr, s->m * 32, cudaMemcpyDeviceToHost, ctx->stream));
  VIMZ_CUDA(cudaMemcpyAsync(Bz, ctx->tmp2.ptr, s->m * 32, cudaMemcpyDeviceToHost, ctx->stream));
  VIMZ_CUDA(cudaMemcpyAsync(Cz, ctx->tmp3.ptr, s->m * 32, cudaMemcpyDeviceToHost, ctx->stream));
  VIMZ_CUDA(cudaStreamSynchronize(ctx->stream));
  return VIMZ_OK;
}

int vimz_commit_T(vimz_ctx* ctx, const vimz_shape* s, const vimz_ck* ck,
                  const vimz_fr* W1, const vimz_fr* u1, const vimz_fr* X1,
                  const vimz_fr* W2, const vimz_fr* X2, vimz_fr* T_out, vimz_point* comm_T) {
  CHECK_ARG(ctx && s && ck && u1 && comm_T && (W1 || s->n == 0) && (W2 || s->n == 0) && ((X1 && X2) || s->io == 0),
            "vimz_commit_T: null argument");
  if (s->m > ck->n) return set_error(VIMZ_ERR_LENGTH, "vimz_commit_T: commitment key shorter than num_cons");
  CtxGuard g(ctx);
  const CurveVTable* vt = curve_vtable(ctx->curve);
  size_t nb = std::max<size_t>(s->n * 32, 32), tb = (1 + s->io) * 32;
  VIMZ_TRY(ctx->tmp0.reserve(nb));
  VIMZ_TRY(ctx->tmp1.reserve(nb));
  VIMZ_TRY(ctx->tmp2.reserve(2 * tb));
  VIMZ_TRY(ctx->tmp3.reserve(std::max<size_t>(s->m * 32, 32)));
  char* t1 = ctx->tmp2.as<char>();
  char* t2 = t1 + tb;
  cudaStream_t st = ctx->stream;
  VIMZ_CUDA(cudaMemcpyAsync(ctx->tmp0.ptr, W1, s->n * 32, cudaMemcpyHostToDevice, st));
  VIMZ_CUDA(cudaMemcpyAsync(ctx->tmp1.ptr, W2, s->n * 32, cudaMemcpyHostToDevice, st));
  VIMZ_CUDA(cudaMemcpyAsync(t1, u1, 32, cudaMemcpyHostToDevice, st));
  VIMZ_CUDA(cudaMemcpyAsync(t2, vt->scalar_one_mont, 32, cudaMemcpyHostToDevice, st));
  if (s->io) {
    VIMZ_CUDA(cudaMemcpyAsync(t1 + 32, X1, s->io * 32, cudaMemcpyHostToDevice, st));
    VIMZ_CUDA(cudaMemcpyAsync(t2 + 32, X2, s->io * 32, cudaMemcpyHostToDevice, st));
  }
  VIMZ_TRY(vt->cross_term(ctx, s, ctx->tmp0.ptr, t1, ctx->tmp1.ptr, t2, ctx->tmp3.ptr, ck, nullptr, nullptr, nullptr));
  if (T_out) VIMZ_CUDA(cudaMemcpyAsync(T_out, ctx->tmp3.ptr, s->m * 32, cudaMemcpyDeviceToHost, st));
  VIMZ_TRY(vt->msm(ctx, 0, ck, 0, ctx->tmp3.ptr, s->m, ctx->ws.result.ptr, s->m > 0));  // no rows: no cross term ran, nothing was recoded
  return fetch_points(ctx, ctx->ws.result.ptr, comm_T, 96);
}

int vimz_fold_witness(vimz_ctx* ctx, const vimz_fr* r, const vimz_fr* W1, const vimz_fr* W2, size_t n,
                      const vimz_fr* E1, const vimz_fr* T, size_t m, vimz_fr* W_out, vimz_fr* E_out) {
  CHECK_ARG(ctx && r && (n == 0 || (W1 && W2 && W_out)) && (m == 0 || (E1 && T && E_out)), "vimz_fold_witness: null argument");
  CtxGuard g(ctx);
  const CurveVTable* vt = curve_vtable(ctx->curve);
  cudaStream_t st = ctx->stream;
  VIMZ_TRY(ctx->tmp0.reserve(std::max<size_t>(n * 32, 32)));
  VIMZ_TRY(ctx->tmp1.reserve(std::max<size_t>(n * 32, 32)));
  VIMZ_TRY(ctx->tmp2.reserve(std::max<size_t>(m * 32, 32)));
  VIMZ_TRY(ctx->tmp3.reserve(std::max<size_t>(m * 32, 32)));
  VIMZ_CUDA(cudaMemcpyAsync(ctx->tmp0.ptr, W1, n * 32, cudaMemcpyHostToDevice, st));
  VIMZ_CUDA(cudaMemcpyAsync(ctx->tmp1.ptr, W2, n * 32, cudaMemcpyHostToDevice, st));
  VIMZ_CUDA(cudaMemcpyAsync(ctx->tmp2.ptr, E1, m * 32, cudaMemcpyHostToDevice, st));
  VIMZ_CUDA(cudaMemcpyAsync(ctx->tmp3.ptr, T, m * 32, cudaMemcpyHostToDevice, st));
  VIMZ_TRY(vt->axpy(ctx, ctx->tmp0.ptr, ctx->tmp1.ptr, r, n, ctx->tmp0.ptr));
  VIMZ_TRY(vt->axpy(ctx, ctx->tmp2.ptr, ctx->tmp3.ptr, r, m, ctx->tmp2.ptr));
  VIMZ_CUDA(cudaMemcpyAsync(W_out, ctx->tmp0.ptr, n * 32, cudaMemcpyDeviceToHost, st));
  VIMZ_CUDA(cudaMemcpyAsync(E_out, ctx->tmp2.ptr, m * 32, cudaMemcpyDeviceToHost, st));
  VIMZ_CUDA(cudaStreamSynchronize(st));
  return VIMZ_OK;
}

// ---- device-resident running instance --------------------------------------------------------------
constexpr size_t ACC_PIN_FRESH = 0, ACC_PIN_COMBINED = 256, ACC_PIN_STAGE = 512;  // layout of vimz_acc::pinned
// vimz_acc::comms, in Jacobian points: the running triple (comm_W, comm_E, K_S), then one (comm_W2, comm_T, P_S) triple per parity
// slot -- step_end folds the three pairs with one launch: running[k] += r * fresh[k]
// ... then the early commitment of a staged witness range and the commitment of the rest (vimz_acc_stage_fresh)
constexpr size_t ACC_SLOT_KS = 2, ACC_SLOT_EARLY = 9, ACC_SLOT_REST = 10, ACC_SLOTS = 11;
static inline char* acc_fresh(const vimz_acc* a, int parity) { return (char*)a->comms + (3 + 3 * parity) * 96; }
void vimz_acc_destroy(vimz_acc* a) {
  if (!a) return;
  vimz_ctx* ctx = a->ctx;
  {
    CtxGuard g(ctx);
    sync_all_streams(ctx);
    void* bufs[] = {a->W1, a->E1, a->W2, a->T, a->tail1, a->tail2, a->comms, a->cache1, a->cache2, a->ksum_scratch, a->ps_parts};
    for (void* b : bufs)
      if (b) cudaFree(b);
    if (a->pinned) cudaFreeHost(a->pinned);
    for (int k = 0; k < 8; k++) {
      if (a->graph[k]) cudaGraphExecDestroy(a->graph[k]);
      if (a->graph_src[k]) cudaGraphDestroy(a->graph_src[k]);
    }
    if (a->ev_stage) cudaEventDestroy(a->ev_stage);
    if (a->ev_auxacc) cudaEventDestroy(a->ev_auxacc);
    if (a->ev_early) cudaEventDestroy(a->ev_early);
    if (a->ev_main) cudaEventDestroy(a->ev_main);
    if (a->ev_w2) cudaEventDestroy(a->ev_w2);
    if (a->ev_aux) cudaEventDestroy(a->ev_aux);
    for (cudaEvent_t ev : {a->ev_ps_fork, a->ev_ps_join, a->ev_ks[0], a->ev_ks[1]})
      if (ev) cudaEventDestroy(ev);
    for (int k = 0; k < 2; k++)
      if (a->ev_side[k]) cudaEventDestroy(a->ev_side[k]);
  }
  // the references taken in acc_create: a key / shape / context destroyed earlier by its owner is freed here
  if (a->ck_w) ck_release(const_cast<vimz_ck*>(a->ck_w));
  if (a->ck) ck_release(const_cast<vimz_ck*>(a->ck));
  if (a->shape) shape_release(const_cast<vimz_shape*>(a->shape));
  delete a;
  ctx_release(ctx);
}

static int acc_create(vimz_ctx* ctx, const vimz_shape* s, const vimz_ck* ck, const vimz_ck* ck_w, size_t w_first, size_t w_count,
                      vimz_acc** out) {
  CtxGuard g(ctx);
  vimz_acc* a = new vimz_acc();
  a->ctx = ctx;
  a->shape = s;
  a->ck = ck;
  a->ck_w = ck_w;
  ctx->refs++;
  const_cast<vimz_shape*>(s)->refs++;
  const_cast<vimz_ck*>(ck)->refs++;
  const_cast<vimz_ck*>(ck_w)->refs++;
  a->w_first = w_first;
  a->w_count = w_count;
  size_t nb = std::max<size_t>(s->n * 32, 32), mb = std::max<size_t>(s->m * 32, 32), tb = (1 + s->io) * 32;
  cudaError_t e = cudaSuccess;
  auto alloc0 = [&](void** p, size_t bytes) {
    if (e != cudaSuccess) return;
    e = cudaMalloc(p, bytes);
    if (e == cudaSuccess) e = cudaMemsetAsync(*p, 0, bytes, ctx->stream);
  };
  alloc0(&a->W1, nb); alloc0(&a->W2, nb); alloc0(&a->E1, mb); alloc0(&a->T, mb);
  alloc0(&a->tail1, tb); alloc0(&a->tail2, tb); alloc0(&a->comms, ACC_SLOTS * 96);
  if (e == cudaSuccess) e = cudaMallocHost(reinterpret_cast<void**>(&a->pinned), ACC_PIN_STAGE + tb);
  if (ctx->opt_cross_cache) {  // the default instance is all zero, and so are its products
    alloc0(&a->cache1, 3 * mb);
    alloc0(&a->cache2, 3 * mb);
  }
  // booleanity-row fold: worth it when a good share of the rows qualifies; needs the cached products (K_S folds like them) and
  // the bucket pipeline (a direct-table key has no final stage that subtracts K_S -- and the shapes it serves have no such rows)
  a->use_ks = ctx->opt_bitrow_fold && a->cache1 && !ck->dtable && s->rowflag && s->n_bitrows * 8 >= s->m && s->n_bitrows >= 32;
  if (a->use_ks) {
    alloc0(&a->ksum_scratch, masked_sum_scratch_bytes(ctx));  // (the arrival counters reset themselves afterwards)
    alloc0(&a->ps_parts, 2 * SCALE_PARTS * 128);
    for (cudaEvent_t* ev : {&a->ev_ps_fork, &a->ev_ps_join, &a->ev_ks[0], &a->ev_ks[1]})
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(ev, cudaEventDisableTiming);
  }
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&a->ev_stage, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&a->ev_auxacc, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&a->ev_early, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&a->ev_main, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&a->ev_w2, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&a->ev_aux, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&a->ev_side[0], cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&a->ev_side[1], cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) {
    vimz_acc_destroy(a);
    return set_error(VIMZ_ERR_CUDA, std::string("vimz_acc_init: ") + cudaGetErrorString(e));
  }
  *out = a;
  return VIMZ_OK;
}

int vimz_acc_init(vimz_ctx* ctx, const vimz_shape* s, const vimz_ck* ck, vimz_acc** out) {
  CHECK_ARG(ctx && s && ck && out, "vimz_acc_init: null argument");
  CHECK_ARG(s->ctx == ctx && ck->ctx == ctx, "vimz_acc_init: shape / key belong to another context");
  if (ck->n < s->m || ck->n < s->n) return set_error(VIMZ_ERR_LENGTH, "vimz_acc_init: commitment key shorter than the shape");
  return acc_create(ctx, s, ck, ck, 0, s->n, out);
}

int vimz_acc_init_sharded(vimz_ctx* ctx, const vimz_shape* s_rows, const vimz_ck* ck_rows, const vimz_ck* ck_vars,
                          size_t var_first, size_t var_count, vimz_acc** out) {
  CHECK_ARG(ctx && s_rows && ck_rows && ck_vars && out, "vimz_acc_init_sharded: null argument");
  CHECK_ARG(s_rows->ctx == ctx && ck_rows->ctx == ctx && ck_vars->ctx == ctx, "vimz_acc_init_sharded: shape / key belong to another context");
  if (var_first > s_rows->n || var_count > s_rows->n - var_first)
    return set_error(VIMZ_ERR_LENGTH, "vimz_acc_init_sharded: variable range outside the witness");
  if (ck_rows->n < s_rows->m || ck_vars->n < var_count)
    return set_error(VIMZ_ERR_LENGTH, "vimz_acc_init_sharded: commitment key shard shorter than its range");
  return acc_create(ctx, s_rows, ck_rows, ck_vars, var_first, var_count, out);
}

// The host side of a fold step waits for ~200 bytes that the stream produces 0.2 - 0.6 ms after the launch: polling the
// stream returns within a few microseconds of completion, a blocking cudaStreamSynchronize adds the wake-up latency of the
// driver's sleeping wait to every step (option "spin_wait", default 1).
static cudaError_t wait_stream(vimz_ctx* ctx, cudaStream_t st) {
  if (!ctx->opt_spin_wait) return cudaStreamSynchronize(st);
  cudaError_t e;
  while ((e = cudaStreamQuery(st)) == cudaErrorNotReady) {
  }
  return e;
}

// main stream waits for the commitment folds still running on the side stream
static int acc_wait_side(vimz_acc* a) {
  for (int k = 0; k < 2; k++)
    if (a->side_pending[k]) {
      VIMZ_CUDA(cudaStreamWaitEvent(a->ctx->stream, a->ev_side[k], 0));
      a->side_pending[k] = false;
    }
  if (a->use_ks)  // (an event that was never recorded counts as complete)
    for (int k = 0; k < 2; k++) VIMZ_CUDA(cudaStreamWaitEvent(a->ctx->stream, a->ev_ks[k], 0));
  return VIMZ_OK;
}

int vimz_acc_load(vimz_acc* a, const vimz_fr* W, const vimz_fr* E, const vimz_fr* u, const vimz_fr* X,
                  const vimz_point* comm_W, const vimz_point* comm_E) {
  CHECK_ARG(a && u && comm_W && comm_E && (W || a->shape->n == 0) && (E || a->shape->m == 0) && (X || a->shape->io == 0),
            "vimz_acc_load: null argument");
  vimz_ctx* ctx = a->ctx;
  CtxGuard g(ctx);
  VIMZ_TRY(acc_wait_side(a));
  cudaStream_t st = ctx->stream;
  const vimz_shape* s = a->shape;
  VIMZ_CUDA(cudaMemcpyAsync(a->W1, W, s->n * 32, cudaMemcpyHostToDevice, st));
  VIMZ_CUDA(cudaMemcpyAsync(a->E1, E, s->m * 32, cudaMemcpyHostToDevice, st));
  VIMZ_CUDA(cudaMemcpyAsync(a->tail1, u, 32, cudaMemcpyHostToDevice, st));
  if (s->io) VIMZ_CUDA(cudaMemcpyAsync((char*)a->tail1 + 32, X, s->io * 32, cudaMemcpyHostToDevice, st));
  VIMZ_CUDA(cudaMemcpyAsync(a->comms, comm_W, 96, cudaMemcpyHostToDevice, st));
  VIMZ_CUDA(cudaMemcpyAsync((char*)a->comms + 96, comm_E, 96, cudaMemcpyHostToDevice, st));
  if (a->cache1 && s->m) {  // (Az1, Bz1, Cz1) of the loaded instance, once
    char* c1 = (char*)a->cache1;
    VIMZ_TRY(curve_vtable(ctx->curve)->spmv3(ctx, s, a->W1, a->tail1, c1, c1 + s->m * 32, c1 + 2 * s->m * 32));
  }
  if (a->use_ks && s->m) {  // K_S of the loaded instance: commit of (A z1) restricted to the booleanity rows, once (general MSM)
    const CurveVTable* vt = curve_vtable(ctx->curve);
    VIMZ_TRY(ctx->tmp5.reserve(std::max<size_t>(s->m * 32, 32)));
    VIMZ_TRY(vt->mask_rows(ctx, st, a->cache1, s->rowflag, s->m, ctx->tmp5.ptr));
    VIMZ_TRY(vt->msm(ctx, 0, a->ck, 0, ctx->tmp5.ptr, s->m, (char*)a->comms + ACC_SLOT_KS * 96, false));
  }
  VIMZ_CUDA(cudaStreamSynchronize(st));
  return VIMZ_OK;
}

// Back to the default relaxed instance (all zero, u = 0), as after vimz_acc_init: a new proof on the same shape / key.
int vimz_acc_reset(vimz_acc* a) {
  CHECK_ARG(a, "vimz_acc_reset: null argument");
  vimz_ctx* ctx = a->ctx;
  CtxGuard g(ctx);
  VIMZ_CUDA(sync_all_streams(ctx));
  a->side_pending[0] = a->side_pending[1] = false;
  const vimz_shape* s = a->shape;
  cudaStream_t st = ctx->stream;
  const size_t nb = std::max<size_t>(s->n * 32, 32), mb = std::max<size_t>(s->m * 32, 32), tb = (1 + s->io) * 32;
  VIMZ_CUDA(cudaMemsetAsync(a->W1, 0, nb, st));
  VIMZ_CUDA(cudaMemsetAsync(a->E1, 0, mb, st));
  VIMZ_CUDA(cudaMemsetAsync(a->tail1, 0, tb, st));
  VIMZ_CUDA(cudaMemsetAsync(a->comms, 0, ACC_SLOTS * 96, st));
  if (a->cache1) VIMZ_CUDA(cudaMemsetAsync(a->cache1, 0, 3 * mb, st));
  VIMZ_CUDA(cudaStreamSynchronize(st));
  a->half_open = false;
  a->step_enqueued = false;
  a->fresh_complete = false;
  return VIMZ_OK;
}

// Cross term + commit(T) on the main lane.  With the booleanity-row fold (vimz_acc::use_ks) the streamed kernels recode the digits
// of T' = T + [bit row] Az1 and the final kernel of the MSM subtracts K_S, so `d_comm_T` receives the true commit(T).  K_S is folded
// by the PREVIOUS step_end on the side stream (its scalar multiplication takes ~0.4 ms and was issued a whole step of the other
// curve ago): the final kernel -- only that one -- waits for that event.
static inline bool acc_shifted(const vimz_acc* a) {
  return a->use_ks && a->ctx->opt_cross_stream && a->shape->n_chunks > 0 && a->shape->m > 0;
}
static int enqueue_cross(vimz_acc* a) {
  const vimz_shape* s = a->shape;
  return curve_vtable(a->ctx->curve)->cross_term(a->ctx, s, a->W1, a->tail1, a->W2, a->tail2, a->T, a->ck, a->cache1, a->cache2,
                                                acc_shifted(a) ? s->rowflag : nullptr);
}
// acc_wait: an event commit(T)'s accumulation kernel waits for (nullptr: none)
static int enqueue_commit_T(vimz_acc* a, void* d_comm_T, void* host_out, cudaEvent_t acc_wait = nullptr) {
  vimz_ctx* ctx = a->ctx;
  const vimz_shape* s = a->shape;
  const CurveVTable* vt = curve_vtable(ctx->curve);
  const bool shifted = acc_shifted(a);
  struct Reset {
    vimz_ctx* c;
    ~Reset() { c->ws.host_out = nullptr; c->ws.sub_jac = nullptr; c->ws.sub_event = nullptr; c->ws.acc_wait = nullptr; }
  } reset{ctx};
  ctx->ws.host_out = host_out;
  ctx->ws.acc_wait = acc_wait;
  if (shifted) {
    ctx->ws.sub_jac = (char*)a->comms + ACC_SLOT_KS * 96;
    ctx->ws.sub_event = a->ev_ks[a->parity ^ 1];
  }
  // (an accumulator without rows -- a shard of a fold spread over more ranks than constraints -- ran no cross term,
  // so nothing recoded T: the commit then does its own, empty, digit pass)
  return vt->msm(ctx, 0, a->ck, 0, a->T, s->m, d_comm_T, s->m > 0);
}
static int enqueue_cross_commit_T(vimz_acc* a, void* d_comm_T, void* host_out) {
  VIMZ_TRY(enqueue_cross(a));
  return enqueue_commit_T(a, d_comm_T, host_out);
}

// P_S = sum over the booleanity rows of (A z2)_i ck_i (what K_S gains, times r, in step_end): a plain sum of the ~55 k bases whose
// fresh wire is 1 -- (A z2)_i is that wire, so only W2 is needed -- into the third slot of the step's fresh triple, then its multiples
// 2^(64 j) P_S (64 doublings by one quad) so that step_end's r * P_S is two half-length chains.  A third branch of the step (stream
// vimz_ctx::ps) forked from `st` where fork_ps_branch was called and complete at ev_ps_join, which the caller makes `st` wait for
// before the step ends: ~0.1 ms + ~0.2 ms of latency beside ~0.5 ms of commitments.  Its nodes are created AFTER those of the main
// lane: in front of them the mat-vec (160 KB of shared memory per SM) waited ~35 us for the SMs to drain this kernel's blocks.
static inline char* acc_ps_parts(const vimz_acc* a, int parity) { return (char*)a->ps_parts + (size_t)parity * SCALE_PARTS * 128; }
static int fork_ps_branch(vimz_acc* a, cudaStream_t st) {  // the point of `st` after which W2 is resident
  if (!(a->use_ks && a->shape->m)) return VIMZ_OK;
  VIMZ_CUDA(cudaEventRecord(a->ev_ps_fork, st));
  return VIMZ_OK;
}
static int enqueue_ps_branch(vimz_acc* a, char* fresh) {
  if (!(a->use_ks && a->shape->m)) return VIMZ_OK;
  vimz_ctx* ctx = a->ctx;
  const CurveVTable* vt = curve_vtable(ctx->curve);
  VIMZ_CUDA(cudaStreamWaitEvent(ctx->ps, a->ev_ps_fork, 0));
  VIMZ_TRY(vt->masked_base_sum(ctx, ctx->ps, a->W2, a->shape->bitcol, a->shape->m, a->ck, a->ksum_scratch, fresh + 2 * 96));
  VIMZ_TRY(vt->point_pow2_parts(ctx, ctx->ps, fresh + 2 * 96, acc_ps_parts(a, a->parity)));
  VIMZ_CUDA(cudaEventRecord(a->ev_ps_join, ctx->ps));
  return VIMZ_OK;
}
static int join_ps_branch(vimz_acc* a, cudaStream_t st) {
  if (!(a->use_ks && a->shape->m)) return VIMZ_OK;
  VIMZ_CUDA(cudaStreamWaitEvent(st, a->ev_ps_join, 0));
  return VIMZ_OK;
}

__global__ void __launch_bounds__(256) k_copy16(uint4* __restrict__ dst, const uint4* __restrict__ src, size_t n16) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
struct CopyNode {
  uint4* dst;
  const uint4* src;
  size_t n16;
  unsigned grid;
};
static CopyNode copy_node_params(const vimz_ctx* ctx, void* dst, const void* src, size_t bytes) {
  CopyNode c;
  c.dst = reinterpret_cast<uint4*>(dst);
  c.src = reinterpret_cast<const uint4*>(src);
  c.n16 = bytes / 16;
  c.grid = (unsigned)std::min<size_t>((c.n16 + 255) / 256, (size_t)ctx->sm_count * 8);
  return c;
}
static int copy_dev(vimz_ctx* ctx, void* dst, const void* src, size_t bytes) {  // bytes: a multiple of 32, both 16-byte aligned
  if (bytes == 0 || dst == src) return VIMZ_OK;
  if ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) {
    VIMZ_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    return VIMZ_OK;
  }
  const size_t n16 = bytes / 16;
  k_copy16<<<(unsigned)std::min<size_t>((n16 + 255) / 256, (size_t)ctx->sm_count * 8), 256, 0, ctx->stream>>>(
      reinterpret_cast<uint4*>(dst), reinterpret_cast<const uint4*>(src), n16);
  VIMZ_LAUNCH_CHECK(ctx);
  return VIMZ_OK;
}

// comm_W2 = commit(ck_w, W2[w_first .. w_first + w_count)) on MSM lane `lane`, result in `fresh` and in the pinned block.  With an early
// commitment of a staged range (vimz_acc_stage_fresh, lane 2, possibly still running: it started a whole step of the other curve
// ago) only the complement is committed here, and one more kernel adds the two -- that one waits for the early lane.
// phase 1 / 2: only the front / back half of the pipeline (MsmWorkspace::phase), the front half recording `acc_record` behind its
// accumulation kernel
static int enqueue_commit_W2(vimz_acc* a, int lane, char* fresh, int phase = 0, cudaEvent_t acc_record = nullptr) {
  vimz_ctx* ctx = a->ctx;
  const CurveVTable* vt = curve_vtable(ctx->curve);
  MsmWorkspace& ws = lane == 0 ? ctx->ws : ctx->ws_aux;
  cudaStream_t st = lane == 0 ? ctx->stream : ctx->aux;
  if (!a->early_valid) {
    ws.host_out = a->pinned + ACC_PIN_FRESH;
    ws.phase = phase;
    ws.acc_record = acc_record;
    int rc = vt->msm(ctx, lane, a->ck_w, 0, (const char*)a->W2 + a->w_first * 32, a->w_count, fresh, lane == 1 && a->w2_counted);
    ws.host_out = nullptr;
    ws.phase = 0;
    ws.acc_record = nullptr;
    return rc;
  }
  // the staged range is a prefix or a suffix of the committed variables (checked by vimz_acc_stage_fresh): the rest is one range
  const size_t rest_first = a->early_first == a->w_first ? a->early_first + a->early_count : a->w_first;
  const size_t rest_count = a->w_count - a->early_count;
  char* comms = (char*)a->comms;
  ws.host_out = nullptr;
  VIMZ_TRY(vt->msm(ctx, lane, a->ck_w, rest_first - a->w_first, (const char*)a->W2 + rest_first * 32, rest_count, comms + ACC_SLOT_REST * 96, false));
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  VIMZ_CUDA(cudaStreamIsCapturing(st, &cs));  // (inside a captured graph an EXTERNAL wait: the early lane is not part of the capture)
  VIMZ_CUDA(cudaStreamWaitEvent(st, a->ev_early, cs == cudaStreamCaptureStatusActive ? cudaEventWaitExternal : cudaEventWaitDefault));
  return vt->point_add2(ctx, st, comms + ACC_SLOT_EARLY * 96, comms + ACC_SLOT_REST * 96, fresh, a->pinned + ACC_PIN_FRESH);
}

// A witness (range) handed over through a host-pointer entry point: an H2D copy -- unless the pointer is device memory (unified
// addressing tells), which the resident-witness callers of the staged entry points pass: then the copy kernel.
static int copy_in(vimz_ctx* ctx, void* dst, const void* src, size_t bytes) {
  if (bytes == 0) return VIMZ_OK;
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, src) == cudaSuccess && at.type == cudaMemoryTypeDevice) return copy_dev(ctx, dst, src, bytes);
  cudaGetLastError();  // (an unregistered host pointer may leave an error behind on old drivers)
  VIMZ_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return VIMZ_OK;
}

// The stream work of step_begin after W2 is resident: everything here has fixed addresses, so it can be captured.
static int enqueue_step_begin(vimz_acc* a, char* fresh) {
  vimz_ctx* ctx = a->ctx;
  const vimz_shape* s = a->shape;
  const CurveVTable* vt = curve_vtable(ctx->curve);
  cudaStream_t st = ctx->stream;
  if (a->w2_mailbox && s->n) {  // step_begin_dev: the resident witness is copied by the step's own first node, whose source
    // parameter is patched before every replay (copy_node_params; reading the address from a mapped host word instead cost 0.8 us per
    // BLOCK: sysmem reads are served one at a time)
    CopyNode cn = copy_node_params(ctx, a->W2, a->w2_src, s->n * 32);
    k_copy16<<<cn.grid, 256, 0, st>>>(cn.dst, cn.src, cn.n16);
    VIMZ_LAUNCH_CHECK(ctx);
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    cudaGraph_t cg = nullptr;
    const cudaGraphNode_t* deps = nullptr;
    size_t ndeps = 0;
    VIMZ_CUDA(cudaStreamGetCaptureInfo_v2(st, &cs, nullptr, &cg, &deps, &ndeps));
    a->cap_copy_node = (cs == cudaStreamCaptureStatusActive && ndeps == 1) ? deps[0] : nullptr;
  }
  // (forked here, with the step: forked behind the cross term it no longer slows the mat-vec by ~15 us, but the aux lane then gets
  // ahead and its accumulation meets the main lane's scatter -- measured +13 us)
  VIMZ_TRY(fork_ps_branch(a, st));
  // tail2 = the staged (1, X2): copied by the kernel that clears commit(T)'s histogram when there is one (one node instead of two)
  const bool fused_tail = ctx->opt_cross_stream && s->n_chunks > 0 && s->m > 0 && !a->ck->dtable;
  if (fused_tail) {
    ctx->ws.pro_src = a->pinned + ACC_PIN_STAGE; ctx->ws.pro_dst = a->tail2; ctx->ws.pro_bytes = (1 + s->io) * 32;
  } else {  // (a kernel that reads the mapped block: a copy-engine node costs more dispatch than it moves)
    const uint32_t nc = (uint32_t)((1 + s->io) * 2);
    k_zero_and_copy<<<(nc + 255) / 256, 256, 0, st>>>(nullptr, 0, reinterpret_cast<const uint4*>(a->pinned + ACC_PIN_STAGE),
                                                      reinterpret_cast<uint4*>(a->tail2), nc);
    VIMZ_LAUNCH_CHECK(ctx);
  }
  // comm_W2 = commit(ck, W2)   (r1cs_instance_and_witness) is independent of T, so it runs on the aux stream with its own
  // workspace while the main stream does the cross term and commit(T).  The main lane is the critical one and is enqueued
  // FIRST: a captured graph dispatches its nodes in creation order, a few microseconds apart, and with the aux lane in
  // front the first cross-term kernel started ~20 us late.
  const bool two_lanes = ctx->opt_aux_lane;
  // Both final kernels also write their result into this accumulator's pinned block (mapped host memory): no D2H copy node.
  struct HostOut {  // cleared on every exit path
    vimz_ctx* c;
    ~HostOut() { c->ws.host_out = nullptr; c->ws_aux.host_out = nullptr; c->ws.pro_bytes = 0; }
  } host_out_guard{ctx};
  a->w2_counted = false;
  if (two_lanes) {
    VIMZ_CUDA(cudaEventRecord(a->ev_w2, st));
    VIMZ_CUDA(cudaStreamWaitEvent(ctx->aux, a->ev_w2, 0));
    // the aux lane's first two nodes (clear + digit pass) are created HERE, ahead of the main lane's: a graph dispatches its root
    // nodes in creation order, and created last the aux lane started ~25 us into the step -- its accumulation then ran beside the
    // main lane's scatter and accumulation instead of beside the (light) cross term
    if (!a->early_valid)
      VIMZ_TRY(vt->msm_digits(ctx, 1, a->ck_w, (const char*)a->W2 + a->w_first * 32, a->w_count, &a->w2_counted));
  } else {  // one lane (used by the profiled pass so kernel times are not inflated by the other lane)
    VIMZ_TRY(enqueue_commit_W2(a, 0, fresh));
  }
  // T = cross term (mat-vecs with z2 + element-wise combination), comm_T = commit(ck, T)      (commit_T)
  // Ordered accumulations (option acc_order): the two accumulation kernels are throughput-bound and gain nothing from sharing the SMs,
  // while everything behind each of them is a latency chain -- so commit(W2)'s (the short one, ready first) runs alone, commit(T)'s
  // starts when it has finished, and the tails of the aux lane are well under way when those of the main lane begin.
  const bool ordered = two_lanes && ctx->opt_acc_order && !a->early_valid && a->w2_counted && !a->ck->dtable && s->m > 0 && !ctx->prof.on;
  VIMZ_TRY(enqueue_cross(a));
  if (ordered) VIMZ_TRY(enqueue_commit_W2(a, 1, fresh, 1, a->ev_auxacc));
  VIMZ_TRY(enqueue_commit_T(a, fresh + 96, a->pinned + ACC_PIN_FRESH + 96, ordered ? a->ev_auxacc : nullptr));
  VIMZ_TRY(enqueue_ps_branch(a, fresh));
  if (two_lanes) {
    VIMZ_TRY(enqueue_commit_W2(a, 1, fresh, ordered ? 2 : 0));
    VIMZ_CUDA(cudaEventRecord(a->ev_aux, ctx->aux));
    VIMZ_CUDA(cudaStreamWaitEvent(st, a->ev_aux, 0));
  }
  return join_ps_branch(a, st);
}

// sync = false: enqueue only (sharded fold: the partial commitments stay on the device for the all-gather)
static int acc_step_begin_common(vimz_acc* a, const vimz_fr* X2, vimz_point* comm_W2, vimz_point* comm_T, bool sync = true) {
  vimz_ctx* ctx = a->ctx;
  const vimz_shape* s = a->shape;
  const CurveVTable* vt = curve_vtable(ctx->curve);
  cudaStream_t st = ctx->stream;
  // comm_W2 / comm_T alternate between two slot pairs so the side-stream fold of the previous
  // step can still read its inputs; the pair used two steps ago must be free again.
  a->parity ^= 1;
  const int p = a->parity;
  if (a->side_pending[p]) {
    VIMZ_CUDA(cudaStreamWaitEvent(st, a->ev_side[p], 0));
    a->side_pending[p] = false;
  }
  char* fresh = acc_fresh(a, p);
  // tail2 = (1, X2) staged in this accumulator's pinned block (read by the copy when the stream reaches it).  A step that
  // was only enqueued (sharded fold) may still have that copy pending: wait for it before overwriting the block.
  if (a->step_enqueued && !a->fresh_complete) VIMZ_CUDA(cudaStreamSynchronize(st));
  uint8_t* stage = a->pinned + ACC_PIN_STAGE;
  memcpy(stage, vt->scalar_one_mont, 32);
  if (s->io) memcpy(stage + 32, X2, s->io * 32);

  // a step with an early commitment / with the witness copy inside has its own launch sequence
  const int gi = p + (a->early_valid ? 2 : 0) + (a->w2_mailbox ? 4 : 0);
  bool use_graph = ctx->opt_graph && !ctx->prof.on && a->warm[gi];
  if (use_graph && a->graph[gi] &&
      (a->graph_epoch[gi] != alloc_epoch() ||  // a workspace moved: rebuild
       (a->early_valid && (a->graph_early[gi][0] != a->early_first || a->graph_early[gi][1] != a->early_count)))) {
    cudaGraphExecDestroy(a->graph[gi]);
    a->graph[gi] = nullptr;
    use_graph = false;  // one eager run re-validates the buffers first
  }
  if (use_graph) {
    if (!a->graph[gi]) {
      uint64_t l0 = ctx->launches, e0 = alloc_epoch();
      VIMZ_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed));
      int rc = enqueue_step_begin(a, fresh);
      cudaGraph_t g = nullptr;
      cudaError_t e = cudaStreamEndCapture(st, &g);
      if (rc != VIMZ_OK || e != cudaSuccess || alloc_epoch() != e0) {
        if (g) cudaGraphDestroy(g);
        cudaGetLastError();
        ctx->opt_graph = false;  // fall back to stream launches for good
        if (rc == VIMZ_OK) rc = enqueue_step_begin(a, fresh);
        if (rc != VIMZ_OK) return rc;
      } else {
        a->graph_launches[gi] = ctx->launches - l0;
        ctx->launches = l0;
        e = cudaGraphInstantiate(&a->graph[gi], g, 0);
        if (a->graph_src[gi]) cudaGraphDestroy(a->graph_src[gi]);
        a->graph_src[gi] = nullptr;
        if (e == cudaSuccess && a->w2_mailbox) a->graph_src[gi] = g;  // the copy node's handle lives in the source graph
        else cudaGraphDestroy(g);
        if (e != cudaSuccess) return set_error(VIMZ_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e));
        a->graph_epoch[gi] = e0;
        a->graph_early[gi][0] = a->early_first; a->graph_early[gi][1] = a->early_count;
        a->copy_node[gi] = a->cap_copy_node;
        if (a->w2_mailbox && !a->copy_node[gi]) return set_error(VIMZ_ERR_CUDA, "step_begin_dev: the captured copy node was not found");
      }
    }
    if (a->graph[gi]) {
      if (a->w2_mailbox && s->n) {  // this replay's witness
        CopyNode cn = copy_node_params(ctx, a->W2, a->w2_src, s->n * 32);
        void* kargs[3] = {&cn.dst, &cn.src, &cn.n16};
        cudaKernelNodeParams kp = {};
        kp.func = reinterpret_cast<void*>(k_copy16);
        kp.gridDim = dim3(cn.grid);
        kp.blockDim = dim3(256);
        kp.kernelParams = kargs;
        VIMZ_CUDA(cudaGraphExecKernelNodeSetParams(a->graph[gi], a->copy_node[gi], &kp));
      }
      VIMZ_CUDA(cudaGraphLaunch(a->graph[gi], st));
      ctx->launches += a->graph_launches[gi];
    }
  } else {
    VIMZ_TRY(enqueue_step_begin(a, fresh));
    a->warm[gi] = true;
  }
  a->early_valid = false;  // consumed
  a->w2_mailbox = false;
  a->fresh_complete = false;
  a->step_enqueued = true;
  if (!sync) return VIMZ_OK;
  VIMZ_CUDA(wait_stream(ctx, st));
  a->fresh_complete = true;
  memcpy(comm_W2, a->pinned + ACC_PIN_FRESH, 96);
  memcpy(comm_T, a->pinned + ACC_PIN_FRESH + 96, 96);
  return VIMZ_OK;
}

int vimz_acc_step_begin_dev_async(vimz_acc* a, const void* d_W2, const vimz_fr* X2, void** d_partials) {
  CHECK_ARG(a && d_partials && (d_W2 || a->shape->n == 0) && (X2 || a->shape->io == 0), "vimz_acc_step_begin_dev_async: null argument");
  CtxGuard g(a->ctx);
  a->early_valid = false;  // the whole witness is replaced: an early commitment of a staged range is void
  if (d_W2 != a->W2)  // (the sharded entry points broadcast the witness straight into the accumulator)
    VIMZ_CUDA(cudaMemcpyAsync(a->W2, d_W2, a->shape->n * 32, cudaMemcpyDeviceToDevice, a->ctx->stream));
  VIMZ_TRY(acc_step_begin_common(a, X2, nullptr, nullptr, false));
  *d_partials = acc_fresh(a, a->parity);
  return VIMZ_OK;
}

int vimz_acc_step_combine_dev(vimz_acc* a, const void* d_gathered, size_t world, vimz_point* comm_W2, vimz_point* comm_T) {
  CHECK_ARG(a && d_gathered && world >= 1 && comm_W2 && comm_T, "vimz_acc_step_combine_dev: null argument");
  vimz_ctx* ctx = a->ctx;
  CtxGuard g(ctx);
  const CurveVTable* vt = curve_vtable(ctx->curve);
  VIMZ_TRY(ctx->ws.result.reserve(2 * 96));
  VIMZ_TRY(vt->point_sum_batch(ctx, d_gathered, world, 2, ctx->ws.result.ptr));
  VIMZ_CUDA(cudaMemcpyAsync(a->pinned + ACC_PIN_COMBINED, ctx->ws.result.ptr, 2 * 96, cudaMemcpyDeviceToHost, ctx->stream));
  VIMZ_CUDA(cudaStreamSynchronize(ctx->stream));
  a->fresh_complete = true;
  memcpy(comm_W2, a->pinned + ACC_PIN_COMBINED, 96);
  memcpy(comm_T, a->pinned + ACC_PIN_COMBINED + 96, 96);
  return VIMZ_OK;
}

int vimz_acc_step_begin(vimz_acc* a, const vimz_fr* W2, const vimz_fr* X2, vimz_point* comm_W2, vimz_point* comm_T) {
  CHECK_ARG(a && comm_W2 && comm_T && (W2 || a->shape->n == 0) && (X2 || a->shape->io == 0), "vimz_acc_step_begin: null argument");
  CtxGuard g(a->ctx);
  a->early_valid = false;
  VIMZ_TRY(copy_in(a->ctx, a->W2, W2, a->shape->n * 32));
  return acc_step_begin_common(a, X2, comm_W2, comm_T);
}

// Streaming upload of the fresh witness.  Most of a step's witness (the Circom part: 118 k of grayscale's 128 k variables) does
// not depend on the previous fold -- the witness generator can run ahead -- so the host may hand it over as soon as the previous
// step_end has been issued: the copy is enqueued on the context's stream behind that step_end (whose axpy is the last reader of
// the buffer) and runs while the host and the GPU work on the OTHER curve.  step_begin_staged then uploads only the rest.
int vimz_acc_stage_fresh(vimz_acc* a, const vimz_fr* W2_part, size_t first, size_t count) {
  CHECK_ARG(a && (W2_part || count == 0), "vimz_acc_stage_fresh: null argument");
  if (first > a->shape->n || count > a->shape->n - first) return set_error(VIMZ_ERR_LENGTH, "vimz_acc_stage_fresh: range outside the witness");
  CtxGuard g(a->ctx);
  if (a->half_open) return set_error(VIMZ_ERR_ARG, "vimz_acc_stage_fresh: a step is open (commit_fresh without cross_begin)");
  vimz_ctx* ctx = a->ctx;
  a->early_valid = false;
  VIMZ_TRY(copy_in(ctx, (char*)a->W2 + first * 32, W2_part, count * 32));
  // Early commitment: the staged range is final, so its share of comm_W2 = commit(ck, W2) can be computed NOW, on a lane of its
  // own, while the GPU folds the other curve (whose step leaves most of the SMs idle) -- the step then commits only the rest and
  // its throughput phase no longer shares the SMs with commit(T).  Only for a prefix / suffix of the committed variables (the rest
  // must be one range), a key on the bucket pipeline, and a range worth a pipeline of its own.
  const bool edge = first == a->w_first || first + count == a->w_first + a->w_count;
  if (ctx->opt_stage_commit && count >= 1024 && first >= a->w_first && first + count <= a->w_first + a->w_count && edge && !a->ck_w->dtable &&
      (a->fresh_complete || !a->step_enqueued)) {
    VIMZ_CUDA(cudaEventRecord(a->ev_stage, ctx->stream));
    VIMZ_CUDA(cudaStreamWaitEvent(ctx->early, a->ev_stage, 0));
    VIMZ_TRY(curve_vtable(ctx->curve)->msm(ctx, 2, a->ck_w, first - a->w_first, (const char*)a->W2 + first * 32, count,
                                          (char*)a->comms + ACC_SLOT_EARLY * 96, false));
    VIMZ_CUDA(cudaEventRecord(a->ev_early, ctx->early));
    a->early_valid = true;
    a->early_first = first;
    a->early_count = count;
  }
  return VIMZ_OK;
}

int vimz_acc_step_begin_staged(vimz_acc* a, const vimz_fr* W2_rest, size_t first, size_t count, const vimz_fr* X2, vimz_point* comm_W2,
                               vimz_point* comm_T) {
  CHECK_ARG(a && ((comm_W2 && comm_T) || (!comm_W2 && !comm_T)) && (W2_rest || count == 0) && (X2 || a->shape->io == 0),
            "vimz_acc_step_begin_staged: null argument");
  if (first > a->shape->n || count > a->shape->n - first) return set_error(VIMZ_ERR_LENGTH, "vimz_acc_step_begin_staged: range outside the witness");
  CtxGuard g(a->ctx);
  if (a->early_valid && count && first < a->early_first + a->early_count && a->early_first < first + count)
    a->early_valid = false;  // the range committed early is being overwritten: commit everything in the step
  VIMZ_TRY(copy_in(a->ctx, (char*)a->W2 + first * 32, W2_rest, count * 32));
  return acc_step_begin_common(a, X2, comm_W2, comm_T, comm_W2 != nullptr);  // (no outputs: enqueue only, vimz_acc_step_wait follows)
}

// step_begin in two halves for a host that has something to enqueue in between (another accumulator's staged upload, its
// own work): _async copies W2 / X2 and enqueues the step, _wait blocks for the two commitments.
int vimz_acc_step_begin_async(vimz_acc* a, const vimz_fr* W2, const vimz_fr* X2) {
  CHECK_ARG(a && (W2 || a->shape->n == 0) && (X2 || a->shape->io == 0), "vimz_acc_step_begin_async: null argument");
  CtxGuard g(a->ctx);
  a->early_valid = false;
  cudaPointerAttributes at;
  if (W2 && cudaPointerGetAttributes(&at, W2) == cudaSuccess && at.type == cudaMemoryTypeDevice && (reinterpret_cast<uintptr_t>(W2) & 15) == 0 &&
      (const void*)W2 != a->W2) {  // a resident witness: its copy is the first node of the step's graph (as in vimz_acc_step_begin_dev)
    a->w2_mailbox = true;
    a->w2_src = W2;
  } else {
    cudaGetLastError();
    VIMZ_TRY(copy_in(a->ctx, a->W2, W2, a->shape->n * 32));
  }
  int rc = acc_step_begin_common(a, X2, nullptr, nullptr, false);
  a->w2_mailbox = false;
  return rc;
}

int vimz_acc_step_wait(vimz_acc* a, vimz_point* comm_W2, vimz_point* comm_T) {
  CHECK_ARG(a && comm_W2 && comm_T, "vimz_acc_step_wait: null argument");
  vimz_ctx* ctx = a->ctx;
  CtxGuard g(ctx);
  if (!a->step_enqueued || a->half_open) return set_error(VIMZ_ERR_ARG, "vimz_acc_step_wait: no enqueued step");
  VIMZ_CUDA(wait_stream(ctx, ctx->stream));
  a->fresh_complete = true;
  memcpy(comm_W2, a->pinned + ACC_PIN_FRESH, 96);
  memcpy(comm_T, a->pinned + ACC_PIN_FRESH + 96, 96);
  return VIMZ_OK;
}

int vimz_acc_step_begin_dev(vimz_acc* a, const void* d_W2, const vimz_fr* X2, vimz_point* comm_W2, vimz_point* comm_T) {
  CHECK_ARG(a && comm_W2 && comm_T && (d_W2 || a->shape->n == 0) && (X2 || a->shape->io == 0), "vimz_acc_step_begin_dev: null argument");
  CtxGuard g(a->ctx);
  // keep W2 resident for step_end (a copy kernel: a copy-engine node in front of the step's graph costs ~10 us of hand-over)
  a->early_valid = false;
  if ((reinterpret_cast<uintptr_t>(d_W2) & 15) == 0 && d_W2 != a->W2) {
    a->w2_mailbox = true;
    a->w2_src = d_W2;
  } else {
    VIMZ_TRY(copy_dev(a->ctx, a->W2, d_W2, a->shape->n * 32));
  }
  int rc = acc_step_begin_common(a, X2, comm_W2, comm_T);
  a->w2_mailbox = false;  // (also on an error path)
  return rc;
}

// ---- the two halves of step_begin as separate calls (strict prove_step order on the secondary curve) --------------
// RecursiveSNARK::prove_step commits the fresh secondary witness at the END of step i (r1cs_instance_and_witness, (6) in
// SURVEY.md section 3.2) and folds it at the START of step i+1 (NIFS::prove, (1)); with host code in between the two
// halves cannot share one call.  commit_fresh opens a step (stages W2 / X2 in the accumulator, returns comm_W2),
// cross_begin finishes what step_begin would have done (T, comm_T); step_end follows as usual.
int vimz_acc_commit_fresh(vimz_acc* a, const vimz_fr* W2, const vimz_fr* X2, vimz_point* comm_W2) {
  CHECK_ARG(a && comm_W2 && (W2 || a->shape->n == 0) && (X2 || a->shape->io == 0), "vimz_acc_commit_fresh: null argument");
  vimz_ctx* ctx = a->ctx;
  CtxGuard g(ctx);
  const vimz_shape* s = a->shape;
  const CurveVTable* vt = curve_vtable(ctx->curve);
  cudaStream_t st = ctx->stream;
  if (a->step_enqueued && !a->fresh_complete) VIMZ_CUDA(cudaStreamSynchronize(st));
  a->early_valid = false;
  a->parity ^= 1;
  const int p = a->parity;
  if (a->side_pending[p]) {
    VIMZ_CUDA(cudaStreamWaitEvent(st, a->ev_side[p], 0));
    a->side_pending[p] = false;
  }
  char* fresh = acc_fresh(a, p);
  uint8_t* stage = a->pinned + ACC_PIN_STAGE;
  memcpy(stage, vt->scalar_one_mont, 32);
  if (s->io) memcpy(stage + 32, X2, s->io * 32);
  VIMZ_CUDA(cudaMemcpyAsync(a->W2, W2, s->n * 32, cudaMemcpyHostToDevice, st));
  VIMZ_CUDA(cudaMemcpyAsync(a->tail2, stage, (1 + s->io) * 32, cudaMemcpyHostToDevice, st));
  VIMZ_TRY(vt->msm(ctx, 0, a->ck_w, 0, (const char*)a->W2 + a->w_first * 32, a->w_count, fresh, false));
  VIMZ_CUDA(cudaMemcpyAsync(a->pinned + ACC_PIN_FRESH, fresh, 96, cudaMemcpyDeviceToHost, st));
  VIMZ_CUDA(cudaStreamSynchronize(st));
  a->step_enqueued = true;
  a->fresh_complete = false;
  a->half_open = true;
  memcpy(comm_W2, a->pinned + ACC_PIN_FRESH, 96);
  return VIMZ_OK;
}

int vimz_acc_cross_begin(vimz_acc* a, vimz_point* comm_T) {
  CHECK_ARG(a && comm_T, "vimz_acc_cross_begin: null argument");
  vimz_ctx* ctx = a->ctx;
  CtxGuard g(ctx);
  if (!a->half_open) return set_error(VIMZ_ERR_ARG, "vimz_acc_cross_begin: no vimz_acc_commit_fresh before it");
  const vimz_shape* s = a->shape;
  const CurveVTable* vt = curve_vtable(ctx->curve);
  cudaStream_t st = ctx->stream;
  char* fresh = acc_fresh(a, a->parity);
  VIMZ_TRY(fork_ps_branch(a, st));
  VIMZ_TRY(enqueue_cross_commit_T(a, fresh + 96, nullptr));
  VIMZ_TRY(enqueue_ps_branch(a, fresh));
  VIMZ_TRY(join_ps_branch(a, st));
  VIMZ_CUDA(cudaMemcpyAsync(a->pinned + ACC_PIN_FRESH + 96, fresh + 96, 96, cudaMemcpyDeviceToHost, st));
  VIMZ_CUDA(cudaStreamSynchronize(st));
  a->half_open = false;
  a->fresh_complete = true;
  memcpy(comm_T, a->pinned + ACC_PIN_FRESH + 96, 96);
  return VIMZ_OK;
}

// The fresh witness staged by the last step_begin / commit_fresh (what nova-snark keeps as l_w_secondary): RecursiveSNARK::verify
// checks it with is_sat before it has been folded.
int vimz_acc_fresh_witness(vimz_acc* a, vimz_fr* W2, vimz_fr* X2) {
  CHECK_ARG(a, "vimz_acc_fresh_witness: null argument");
  vimz_ctx* ctx = a->ctx;
  CtxGuard g(ctx);
  const vimz_shape* s = a->shape;
  if (W2 && s->n) VIMZ_CUDA(cudaMemcpyAsync(W2, a->W2, s->n * 32, cudaMemcpyDeviceToHost, ctx->stream));
  if (X2 && s->io) VIMZ_CUDA(cudaMemcpyAsync(X2, (char*)a->tail2 + 32, s->io * 32, cudaMemcpyDeviceToHost, ctx->stream));
  VIMZ_CUDA(cudaStreamSynchronize(ctx->stream));
  return VIMZ_OK;
}

int vimz_acc_step_end(vimz_acc* a, const vimz_fr* r) {
  CHECK_ARG(a && r, "vimz_acc_step_end: null argument");
  vimz_ctx* ctx = a->ctx;
  CtxGuard g(ctx);
  if (a->half_open) return set_error(VIMZ_ERR_ARG, "vimz_acc_step_end: vimz_acc_commit_fresh without vimz_acc_cross_begin");
  if (!a->step_enqueued) return set_error(VIMZ_ERR_ARG, "vimz_acc_step_end: no step_begin before it");
  const vimz_shape* s = a->shape;
  const CurveVTable* vt = curve_vtable(ctx->curve);
  cudaStream_t st = ctx->stream;
  char* comms = (char*)a->comms;
  char* fresh = acc_fresh(a, a->parity);
  // r travels by value in both launches: W1 += r*W2, E1 += r*T, (u1, X1) += r*(1, X2) on the main stream ...
  AxpySeg segs[4] = {{a->W1, a->W2, s->n}, {a->E1, a->T, s->m}, {a->tail1, a->tail2, 1 + s->io}, {a->cache1, a->cache2, a->cache1 ? 3 * s->m : 0}};
  VIMZ_TRY(vt->axpyn(ctx, segs, 4, r));
  // ... and comm_W1 += r*comm_W2 ; comm_E1 += r*comm_T on the side stream: two 128-bit scalar multiplications,
  // latency-bound, overlapping the next step's MSMs.  Their inputs (the step's fresh pair) were complete when
  // step_begin returned, and the running pair is only touched on the side stream, so no cross-stream wait is needed
  // (unless the step was only enqueued and never waited for: then the side stream waits for the main stream).
  const bool ks = a->use_ks && s->m;
  if (!a->fresh_complete) {
    VIMZ_CUDA(cudaEventRecord(a->ev_main, st));
    VIMZ_CUDA(cudaStreamWaitEvent(ctx->side, a->ev_main, 0));
    if (ks) VIMZ_CUDA(cudaStreamWaitEvent(ctx->ks, a->ev_main, 0));
  }
  // With the booleanity-row fold K_S += r * P_S goes by itself on its own stream: the next step of this accumulator subtracts K_S in
  // the final kernel of its commit(T), ~0.6 ms from now, while the lone warp of a 128-bit scalar multiplication takes 0.4 - 0.65 ms
  // beside busy SMs (and behind the other two on one stream the folds of a step took longer than the step).  step_begin left the
  // multiples 2^(64 j) P_S, so this one is two 64-bit pieces (~0.25 ms).
  if (ks) {
    VIMZ_TRY(vt->point_scale_add_parts(ctx, ctx->ks, comms + ACC_SLOT_KS * 96, r, acc_ps_parts(a, a->parity), comms + ACC_SLOT_KS * 96));
    VIMZ_CUDA(cudaEventRecord(a->ev_ks[a->parity], ctx->ks));
  }
  VIMZ_TRY(vt->point_scale_add_val(ctx, ctx->side, comms, r, fresh, comms, 2));
  VIMZ_CUDA(cudaEventRecord(a->ev_side[a->parity], ctx->side));
  a->side_pending[a->parity] = true;
  return VIMZ_OK;
}

int vimz_acc_download(vimz_acc* a, vimz_fr* W, vimz_fr* E, vimz_fr* u, vimz_fr* X, vimz_point* comm_W, vimz_point* comm_E) {
  CHECK_ARG(a, "vimz_acc_download: null argument");
  vimz_ctx* ctx = a->ctx;
  CtxGuard g(ctx);
  VIMZ_TRY(acc_wait_side(a));
  cudaStream_t st = ctx->stream;
  const vimz_shape* s = a->shape;
  if (W) VIMZ_CUDA(cudaMemcpyAsync(W, a->W1, s->n * 32, cudaMemcpyDeviceToHost, st));
  if (E) VIMZ_CUDA(cudaMemcpyAsync(E, a->E1, s->m * 32, cudaMemcpyDeviceToHost, st));
  if (u) VIMZ_CUDA(cudaMemcpyAsync(u, a->tail1, 32, cudaMemcpyDeviceToHost, st));
  if (X && s->io) VIMZ_CUDA(cudaMemcpyAsync(X, (char*)a->tail1 + 32, s->io * 32, cudaMemcpyDeviceToHost, st));
  if (comm_W) VIMZ_CUDA(cudaMemcpyAsync(comm_W, a->comms, 96, cudaMemcpyDeviceToHost, st));
  if (comm_E) VIMZ_CUDA(cudaMemcpyAsync(comm_E, (char*)a->comms + 96, 96, cudaMemcpyDeviceToHost, st));
  VIMZ_CUDA(cudaStreamSynchronize(st));
  return VIMZ_OK;
}

int vimz_acc_last_T(vimz_acc* a, vimz_fr* T) {
  CHECK_ARG(a && (T || a->shape->m == 0), "vimz_acc_last_T: null argument");
  CtxGuard g(a->ctx);
  VIMZ_CUDA(cudaMemcpyAsync(T, a->T, a->shape->m * 32, cudaMemcpyDeviceToHost, a->ctx->stream));
  VIMZ_CUDA(cudaStreamSynchronize(a->ctx->stream));
  return VIMZ_OK;
}

// ---- test / bench utilities --------------------------------------------------------------------------
int vimz_gen_bases_dev(vimz_ctx* ctx, uint64_t k0, uint64_t dk, size_t n, void* d_out) {
  CHECK_ARG(ctx && (d_out || n == 0), "vimz_gen_bases_dev: null argument");
  CtxGuard g(ctx);
  VIMZ_TRY(curve_vtable(ctx->curve)->gen_bases(ctx, k0, dk, n, d_out));
  VIMZ_CUDA(cudaStreamSynchronize(ctx->stream));
  return VIMZ_OK;
}

int vimz_field_op(vimz_ctx* ctx, int which, int op, const vimz_fr* a, const vimz_fr* b, size_t n, vimz_fr* out) {
  CHECK_ARG(ctx && (n == 0 || (a && b && out)) && which >= 0 && which <= 1 && op >= 0 && op <= 3, "vimz_field_op: bad argument");
  CtxGuard g(ctx);
  size_t bytes = std::max<size_t>(n * 32, 32);
  VIMZ_TRY(ctx->tmp0.reserve(bytes));
  VIMZ_TRY(ctx->tmp1.reserve(bytes));
  VIMZ_TRY(ctx->tmp2.reserve(bytes));
  VIMZ_CUDA(cudaMemcpyAsync(ctx->tmp0.ptr, a, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  VIMZ_CUDA(cudaMemcpyAsync(ctx->tmp1.ptr, b, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  VIMZ_TRY(curve_vtable(ctx->curve)->field_op(ctx, which, op, ctx->tmp0.ptr, ctx->tmp1.ptr, n, ctx->tmp2.ptr));
  VIMZ_CUDA(cudaMemcpyAsync(out, ctx->tmp2.ptr, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
  VIMZ_CUDA(cudaStreamSynchronize(ctx->stream));
  return VIMZ_OK;
}

}  // extern "C"
