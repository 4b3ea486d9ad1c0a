// common.cuh -- context, error handling and device buffers shared by the kernels' host drivers.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <atomic>
#include <cstdio>
#include <mutex>
#include <string>
#include <vector>
#include "../../include/vimz_gpu.h"

namespace vimz {

extern thread_local std::string g_last_error;

inline int set_error(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

#define VIMZ_CUDA(expr)                                                                            \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) {                                                                       \
      char _buf[512];                                                                              \
      snprintf(_buf, sizeof(_buf), "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return ::vimz::set_error(VIMZ_ERR_CUDA, _buf);                                               \
    }                                                                                              \
  } while (0)

#define VIMZ_TRY(expr)             \
  do {                             \
    int _rc = (expr);              \
    if (_rc != VIMZ_OK) return _rc; \
  } while (0)

// Bumped whenever a DevBuf is (re)allocated: captured CUDA graphs hold raw pointers and are rebuilt when it moves.
inline std::atomic<uint64_t>& alloc_epoch() {
  static std::atomic<uint64_t> e{0};
  return e;
}

// Grow-only device buffer (no per-call cudaMalloc on the hot path).
struct DevBuf {
  void* ptr = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes) {
    if (bytes <= cap) return VIMZ_OK;
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    VIMZ_CUDA(cudaMalloc(&ptr, want));
    cap = want;
    alloc_epoch()++;
    return VIMZ_OK;
  }
  void release() {
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    cap = 0;
  }
  template <class T>
  T* as() const { return reinterpret_cast<T*>(ptr); }
};

}  // namespace vimz

// MSM scratch: everything Pippenger needs besides the resident table.
struct MsmWorkspace {
  vimz::DevBuf counts;     // [M]   entries per bucket
  vimz::DevBuf offsets;    // [M+1] exclusive prefix of counts
  vimz::DevBuf cursor;     // [M]   scatter cursors
  vimz::DevBuf blocksums;  // scan scratch
  vimz::DevBuf digits;     // [nwin][n] recoded digits (magnitude | sign<<31, 0 = none), written once per MSM
  vimz::DevBuf sorted;     // [E]   table index | sign<<31, grouped by bucket
  vimz::DevBuf cls;        // control words (giant-bucket counter)
  vimz::DevBuf biglist;    // ids of buckets cut into many segments
  vimz::DevBuf partials;   // [2 * nthreads] XYZZ head/tail partial sums of the accumulation segments
  vimz::DevBuf buckets;    // [M] XYZZ
  vimz::DevBuf chunkA;     // [T] XYZZ chunk sums
  vimz::DevBuf chunkL;     // [T] XYZZ chunk weighted sums
  vimz::DevBuf bitsums;    // [(nb+1) * G] XYZZ
  vimz::DevBuf scaled;     // [nb+1] XYZZ
  vimz::DevBuf deferred;   // [max_giants] XYZZ: weighted sums of the giant buckets kept out of the bucket array
  vimz::DevBuf scal;       // staged scalars (host-pointer entry points)
  vimz::DevBuf result;     // Jacobian results (device)
  // set by a caller around ONE msm launch sequence: the final kernel also writes its Jacobian result here -- mapped page-locked
  // host memory the host reads after the stream completes (saves the D2H copy node of a fold step and its dispatch gap)
  void* host_out = nullptr;
  // same protocol: a Jacobian point the final kernel subtracts from the sum (the accumulator's K_S), and the event after which
  // it is valid (waited for right in front of the final kernel only)
  const void* sub_jac = nullptr;
  cudaEvent_t sub_event = nullptr;
  // same protocol: a small host -> device copy (the accumulator's staged (1, X2) tail, in mapped page-locked memory) that the kernel
  // clearing the control block + histogram performs as well -- one graph node instead of a copy node and a memset node in front of
  // the first kernel of the critical lane (each costs ~8 us of dispatch there)
  // same protocol: run only the FRONT half of the bucket pipeline (1: clear / digits / scan / scatter / accumulate) or only the BACK
  // half (2: combine / reduce), so that a caller can create the nodes of two lanes in an order of its choice; and events around the
  // accumulation kernel -- waited for right in front of it / recorded right behind it -- to order the accumulations of two lanes
  int phase = 0;
  cudaEvent_t acc_wait = nullptr, acc_record = nullptr;
  const void* pro_src = nullptr;
  void* pro_dst = nullptr;
  size_t pro_bytes = 0;    // multiple of 16
  uint32_t last_M = 0;     // buckets of the last MSM run on this workspace (statistics: vimz_ctx_profile "laneK_*")
  void release() {
    counts.release(); offsets.release(); cursor.release(); blocksums.release(); sorted.release(); digits.release();
    cls.release(); biglist.release(); partials.release(); buckets.release();
    chunkA.release(); chunkL.release(); bitsums.release(); scaled.release(); deferred.release(); scal.release(); result.release();
  }
};

// Optional per-phase device timers (vimz_ctx_set_option("profile", 1)): event pairs are recorded on the
// context stream around a phase and resolved when vimz_ctx_profile() is called.
enum ProfTimer { PROF_MSM_SORT = 0, PROF_MSM_ACCUMULATE, PROF_MSM_REDUCE, PROF_CROSS_TERM, PROF_AXPY, PROF_SPMV, PROF_MSM_ACC_KERNEL,
                 PROF_MSM_ACC_KERNEL_FUSED,  // the accumulation launches whose digits came from the fused cross term: commit(T)
                 PROF_COUNT };
struct ProfSpan { cudaEvent_t a, b; int timer; };
struct Profiler {
  bool on = false;
  std::vector<ProfSpan> open;          // recorded, not yet resolved
  std::vector<cudaEvent_t> pool;       // reusable events
  double ms[PROF_COUNT] = {0};
  uint64_t calls[PROF_COUNT] = {0};
  uint64_t msm_entries = 0;            // bucket insertions (non-zero digits) seen by profiled MSMs
  uint64_t msm_entries_fused = 0;      // ... of which in commit(T) launches (digits recoded by the cross term)
  std::vector<uint32_t*> entry_slots;  // pinned words receiving each MSM's entry total
  std::vector<bool> entry_fused;       // parallel to entry_slots
  std::vector<uint32_t*> entry_pool;
};

struct vimz_ctx {
  // Every entry point that touches the context's streams or workspaces holds this lock for the call: nova-snark's
  // prove_step is sequential, but CompressedSNARK::prove / RecursiveSNARK::verify commit from several rayon workers at
  // once, and those calls must not interleave on one workspace.  Recursive: entry points nest.
  std::recursive_mutex mu;
  // 1 for the handle returned by vimz_ctx_create + 1 per live key / shape / accumulator: vimz_ctx_destroy only drops
  // the owner's reference, the context is freed when the last child is destroyed (see capi.cu "handle lifetimes")
  std::atomic<int> refs{1};
  int curve = 0;
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t side = nullptr;  // instance-fold scalar multiplications overlap the next step here
  cudaStream_t aux = nullptr;   // second MSM lane: commit(W2) runs beside cross-term + commit(T)
  cudaStream_t ps = nullptr;    // third branch of a step: P_S of the booleanity-row fold and its 2^(64 j) multiples (vimz_acc::use_ks)
  cudaStream_t early = nullptr; // lane 2: the commitment of a witness range staged ahead of its step (vimz_acc_stage_fresh), beside the OTHER curve's step
  cudaStream_t ks = nullptr;    // K_S += r * P_S of step_end: a lone warp with a deadline (the next commit(T) of the accumulator), kept apart
                                // from the comm_W / comm_E folds on `side`, which have none
  int sm_count = 148;
  long opt_window = 0;  // 0 = auto
  bool opt_graph = true; // replay the fixed launch sequence of a fold step as a CUDA graph
  bool opt_defer_giants = true; // giant buckets are summed beside the bucket reduction (k_reduce_tail) instead of in front of it
  bool opt_acc_order = false;   // fold step: commit(T)'s accumulation starts when commit(W2)'s has finished instead of sharing the SMs with it (measured: +33 us per step)
  bool opt_stage_commit = false; // vimz_acc_stage_fresh also commits the staged range at once (lane 2); step_begin_staged commits only the rest
  bool opt_bitrow_fold = true;  // accumulators created from now on keep K_S and commit T + [bit row] Az1 (r1cs.cuh, k_cross_finish)
  bool opt_spin_wait = true; // step_begin polls the stream for its result instead of a blocking synchronise
  long opt_acc_blocks = 4; // 128-thread accumulation blocks per SM (4 = register-file limit)
  long opt_seg_min_aux = 0; // the same for lane 1 (0 = opt_seg_min)
  long opt_seg_min = 8;    // shortest accumulation segment (entries per thread): fewer => more threads busy on small MSMs
  bool opt_cross_stream = true; // cross term: chunked CSR streaming through shared memory (false: row-class kernel)
  bool opt_cross_cache = true;  // accumulators created from now on keep (Az1, Bz1, Cz1) resident instead of recomputing them
  bool opt_aux_lane = true; // fold step: commit(W2) on the aux stream beside cross term + commit(T)
  long opt_direct_c = 0;       // digit width of the direct table (0 = chosen by key length)
  long opt_direct_bps = 2;     // k_msm_direct blocks per SM (1..4): at 2 the two commits of a fold step (two stream lanes) are resident together
  long opt_direct_max = 32768; // keys up to this many points get the direct multiples table (256 KB per point); 0 = never
  uint64_t launches = 0;
  MsmWorkspace ws, ws_aux, ws_early;  // lanes 0 / 1 / 2 (vimz_ctx::stream / aux / early)
  vimz::DevBuf tmp0, tmp1, tmp2, tmp3, tmp4, tmp5;  // R1CS staging for host-pointer entry points
  void* pinned = nullptr;                      // small pinned staging block for results
  Profiler prof;
};

namespace vimz {
struct ProfScope {
  vimz_ctx* ctx;
  ProfSpan span;
  bool active;
  cudaStream_t st;
  ProfScope(vimz_ctx* c, int timer, cudaStream_t stream) : ctx(c), active(c && c->prof.on), st(stream) {
    if (!active) return;
    auto get = [&]() {
      cudaEvent_t e;
      if (!ctx->prof.pool.empty()) { e = ctx->prof.pool.back(); ctx->prof.pool.pop_back(); }
      else cudaEventCreate(&e);
      return e;
    };
    span.a = get(); span.b = get(); span.timer = timer;
    cudaEventRecord(span.a, st);
  }
  ~ProfScope() {
    if (!active) return;
    cudaEventRecord(span.b, st);
    ctx->prof.open.push_back(span);
  }
};
}  // namespace vimz

struct vimz_ck {
  vimz_ctx* ctx = nullptr;
  std::atomic<int> refs{1};  // owner handle + accumulators committing with this key
  size_t n = 0;
  int c = 0;          // window bits
  int nwin = 0;       // windows
  void* table = nullptr;  // [nwin][n] affine points, row j = 2^(c*j) * base
  // short keys (n <= option "msm_direct_max"): every digit multiple k * table[j][i], k = 1 .. 2^(c-1), so the MSM is a
  // plain sum with no buckets (msm.cuh, k_msm_direct); nullptr = bucket pipeline
  void* dtable = nullptr;
};

struct vimz_shape {
  vimz_ctx* ctx = nullptr;
  std::atomic<int> refs{1};  // owner handle + accumulators folding over this shape
  size_t m = 0, n = 0, io = 0;
  size_t nnz[3] = {0, 0, 0};
  uint32_t* rowptr[3] = {nullptr, nullptr, nullptr};  // [m+1]
  uint32_t* col[3] = {nullptr, nullptr, nullptr};     // [nnz]
  void* val[3] = {nullptr, nullptr, nullptr};         // [nnz] Montgomery scalars
  // streamed mat-vec (r1cs.cuh, k_matvec_stream): rows cut into chunks of bounded non-zeros; every chunk's (column,
  // coefficient-dictionary index) pairs packed contiguously (A's, then B's, then C's) and 16-byte aligned for the TMA bulk copy
  void* dict = nullptr;                               // [n_dict] distinct Montgomery coefficients of A, B, C (0: +1, 1: -1, no multiplication)
  size_t n_dict = 0;
  void* chunk_stream = nullptr;                       // uint2 (col, vidx) pairs
  void* chunk_desc = nullptr;                         // [n_chunks] ChunkDesc
  size_t n_chunks = 0;
  uint32_t* long_rows = nullptr;                      // rows with > R1CS_LONG_ROW non-zeros over A+B+C (warp each)
  size_t n_long = 0;
  uint32_t* mid_rows = nullptr;                       // rows with R1CS_SHORT_ROW+1 .. R1CS_LONG_ROW non-zeros (8 lanes each)
  size_t n_mid = 0;
  uint8_t* rowflag = nullptr;                         // [m] 1 = booleanity row b*(b-1)=0 (A = {(b,1)}, B = {(b,1),(one,-1)}, C = {}), see k_cross_finish
  uint32_t* bitcol = nullptr;                         // [m] the column b of such a row (a variable of W), ~0 for every other row
  size_t n_bitrows = 0;
};

struct vimz_acc {
  vimz_ctx* ctx = nullptr;
  const vimz_shape* shape = nullptr;
  const vimz_ck* ck = nullptr;    // table over the rows of this accumulator's shape (commit(T), commit(E))
  // commit(W2) = msm(ck_w, W2[w_first .. w_first + w_count)).  Unsharded: ck_w = ck, the whole witness.  Row-range
  // shard (vimz_acc_init_sharded): ck_w covers ck[w_first ..), W stays replicated, E / T hold only the local rows.
  const vimz_ck* ck_w = nullptr;
  size_t w_first = 0, w_count = 0;
  void *W1 = nullptr, *E1 = nullptr, *W2 = nullptr, *T = nullptr;
  void *tail1 = nullptr, *tail2 = nullptr;  // [1+io]: (u, X)
  // cached products (option "cross_cache", fixed when the accumulator is created): cache1 = (Az1, Bz1, Cz1)[3][m] of the
  // running instance, folded in step_end with cache2 = (Az2, Bz2, Cz2)[3][m] of the fresh one -- linear, so exact
  void *cache1 = nullptr, *cache2 = nullptr;
  // Jacobian points comm_W1, comm_E1, K_S, then two alternating (comm_W2, comm_T, P_S) triples (ACC_SLOTS, capi.cu)
  void* comms = nullptr;
  // Booleanity-row fold (r1cs.cuh, k_cross_finish): K_S = sum_{i in S} (A z1)_i ck_i of the running instance lives in the third
  // commitment slot and folds with the other two (K_S += r * P_S, P_S = the same sum for the step's fresh z2, made by
  // k_masked_base_sum on the side stream).  Fixed when the accumulator is created.
  bool use_ks = false;
  void* ksum_scratch = nullptr;  // block / group sums + self-resetting arrival counters of k_masked_base_sum
  void* ps_parts = nullptr;      // [2 parities][SCALE_PARTS] XYZZ records 2^(32 j) P_S (k_point_pow2_parts), read by step_end
  cudaEvent_t ev_ps_fork = nullptr, ev_ps_join = nullptr;  // the P_S branch of a step (stream vimz_ctx::ps)
  cudaEvent_t ev_ks[2] = {nullptr, nullptr};               // "K_S has absorbed the step of this parity" (side stream)
  // pinned host block of THIS accumulator (several accumulators may share a context and have steps in flight at once):
  // [0, 192) the step's (comm_W2, comm_T) as copied back by the stream, [256, 448) the combined pair of a sharded step,
  // [512, ..) the staged (1, X2) tail that the step's H2D copy reads when the stream reaches it
  uint8_t* pinned = nullptr;
  cudaEvent_t ev_main = nullptr, ev_side[2] = {nullptr, nullptr}, ev_w2 = nullptr, ev_aux = nullptr;
  bool side_pending[2] = {false, false};
  bool fresh_complete = false;  // the host has waited for the last step_begin (sync entry points, step_combine_dev)
  bool step_enqueued = false;   // a step_begin has been issued on this accumulator
  bool half_open = false;       // vimz_acc_commit_fresh done, vimz_acc_cross_begin still to come
  int parity = 0;
  // step_begin's launch sequence (cross term + both MSMs, ~35 kernels on two streams) captured once per
  // parity slot and replayed: the step is latency-bound and stream launches cost more than the small kernels
  // (index = parity + 2 * [the step has an early commitment of a staged range: it commits only the rest of W2] + 4 * [the copy
  // of a resident witness is the first node])
  cudaGraphExec_t graph[8] = {};
  uint64_t graph_epoch[8] = {};
  uint64_t graph_launches[8] = {};
  bool warm[8] = {};
  size_t graph_early[8][2] = {};  // the staged range a graph with an early commitment was captured for
  // Early commitment (option "stage_commit"): vimz_acc_stage_fresh committed W2[early_first .. + early_count) on lane 2 into the
  // EARLY slot of `comms`; the next step_begin_staged commits the complement and adds the two.  Consumed by that step.
  bool w2_counted = false;  // inside enqueue_step_begin: the aux lane's digit pass was created ahead of the main lane's nodes
  // W2 of a step_begin_dev call: the copy of the resident witness is the FIRST NODE of the step's graph, its source parameter patched
  // before every replay (a separate copy launch in front of the graph costs ~10 us of dispatch)
  bool w2_mailbox = false;
  const void* w2_src = nullptr;
  cudaGraphNode_t cap_copy_node = nullptr, copy_node[8] = {};
  cudaGraph_t graph_src[8] = {};  // kept alive for the graphs that have a copy node (its handle belongs to the source graph)
  bool early_valid = false;
  size_t early_first = 0, early_count = 0;
  cudaEvent_t ev_stage = nullptr, ev_early = nullptr;
  cudaEvent_t ev_auxacc = nullptr;  // "commit(W2)'s accumulation kernel has finished" inside a step (option acc_order)
};
