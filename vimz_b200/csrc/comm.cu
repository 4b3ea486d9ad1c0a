// comm.cu -- multi-GPU entry points of libvimz_gpu.so: one fold step / one MSM spread over the GPUs of a node, with the
// exchange done INSIDE the library (NCCL on the context's own stream), so a non-Python host -- the Rust prover reached from
// /root/reference/vimz/src/nova_snark_backend/mod.rs:19-20 -- drives 8 GPUs through the same C ABI as one.
//
// What is exchanged: elliptic-curve addition is not an NCCL reduction, so every rank contributes its
// partial commitments (96 bytes each) to an all-gather and adds the `world` partials itself (k_point_sum_batch): every rank
// ends up with the same full commitment and derives the same Fiat-Shamir challenge.  The fresh witness, when it lives on
// one rank's host, is broadcast over NVLink after that rank's H2D copy.
//
// NCCL is bound at run time (dlopen of libnccl.so.2 on the first vimz_comm_* call): the library has no link-time
// dependency on it, single-GPU users never load it, and inside a torch process the already loaded NCCL is reused.
#include <dlfcn.h>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include "curve_impl.cuh"

using namespace vimz;

// ---- minimal NCCL ABI (stable across NCCL 2.x) -------------------------------------------------------------------------
namespace {
struct NcclUniqueId { char internal[128]; };
typedef struct ncclComm* NcclComm;
enum { NCCL_SUCCESS = 0 };
enum { NCCL_INT8 = 0, NCCL_UINT8 = 1 };

struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
  int (*Broadcast)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*GetVersion)(int*) = nullptr;
  std::string error;
};

NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {getenv("VIMZ_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      if (!n || !*n) continue;
      api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (!api.handle) {
      api.error = "libnccl.so.2 not found (set VIMZ_NCCL_LIB): the multi-GPU entry points need NCCL";
      return;
    }
    auto sym = [&](const char* s) {
      void* p = dlsym(api.handle, s);
      if (!p && api.error.empty()) api.error = std::string("NCCL symbol missing: ") + s;
      return p;
    };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
    api.Broadcast = reinterpret_cast<decltype(api.Broadcast)>(sym("ncclBroadcast"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(sym("ncclGetVersion"));
  });
  return api;
}

int nccl_fail(const char* what, int rc) {
  NcclApi& a = nccl();
  return set_error(VIMZ_ERR_CUDA, std::string(what) + ": " + (a.GetErrorString ? a.GetErrorString(rc) : "NCCL error"));
}
#define VIMZ_NCCL(expr)                                   \
  do {                                                    \
    int _r = (expr);                                      \
    if (_r != NCCL_SUCCESS) return nccl_fail(#expr, _r);  \
  } while (0)

struct DevGuard {
  int prev = -1, dev;
  explicit DevGuard(int d) : dev(d) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) cudaSetDevice(dev);
  }
  ~DevGuard() {
    if (prev >= 0 && prev != dev) cudaSetDevice(prev);
  }
};
}  // namespace

struct vimz_comm {
  NcclComm comm = nullptr;
  int rank = 0, world = 1, device = 0;
  void* d_gather = nullptr;  // [world][2] Jacobian points: the all-gathered partial commitments of one step
  std::mutex mu;
};

extern "C" {

int vimz_comm_unique_id(uint8_t id[VIMZ_COMM_ID_BYTES]) {
  if (!id) return set_error(VIMZ_ERR_ARG, "vimz_comm_unique_id: null argument");
  NcclApi& a = nccl();
  if (!a.error.empty()) return set_error(VIMZ_ERR_NO_DEVICE, a.error);
  NcclUniqueId u;
  VIMZ_NCCL(a.GetUniqueId(&u));
  static_assert(sizeof(NcclUniqueId) == VIMZ_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
  memcpy(id, &u, sizeof(u));
  return VIMZ_OK;
}

int vimz_comm_create(int device, const uint8_t id[VIMZ_COMM_ID_BYTES], int rank, int world, vimz_comm** out) {
  if (!out || !id) return set_error(VIMZ_ERR_ARG, "vimz_comm_create: null argument");
  *out = nullptr;
  if (world < 1 || rank < 0 || rank >= world) return set_error(VIMZ_ERR_ARG, "vimz_comm_create: bad rank / world");
  int ndev = vimz_device_count();
  if (ndev <= 0) return set_error(VIMZ_ERR_NO_DEVICE, "vimz_comm_create: no CUDA device visible");
  if (device < 0 || device >= ndev) return set_error(VIMZ_ERR_ARG, "vimz_comm_create: device index out of range");
  NcclApi& a = nccl();
  if (!a.error.empty()) return set_error(VIMZ_ERR_NO_DEVICE, a.error);
  DevGuard dg(device);
  vimz_comm* c = new vimz_comm();
  c->rank = rank; c->world = world; c->device = device;
  NcclUniqueId u;
  memcpy(&u, id, sizeof(u));
  int rc = a.CommInitRank(&c->comm, world, u, rank);
  if (rc != NCCL_SUCCESS) {
    delete c;
    return nccl_fail("ncclCommInitRank", rc);
  }
  cudaError_t e = cudaMalloc(&c->d_gather, (size_t)world * 2 * 96);
  if (e != cudaSuccess) {
    a.CommDestroy(c->comm);
    delete c;
    return set_error(VIMZ_ERR_CUDA, std::string("vimz_comm_create: ") + cudaGetErrorString(e));
  }
  *out = c;
  return VIMZ_OK;
}

void vimz_comm_destroy(vimz_comm* c) {
  if (!c) return;
  DevGuard dg(c->device);
  cudaDeviceSynchronize();
  if (c->d_gather) cudaFree(c->d_gather);
  if (c->comm) nccl().CommDestroy(c->comm);
  delete c;
}

int vimz_comm_rank(const vimz_comm* c) { return c ? c->rank : -1; }
int vimz_comm_world(const vimz_comm* c) { return c ? c->world : 0; }
int vimz_comm_nccl_version(void) {
  NcclApi& a = nccl();
  int v = 0;
  if (a.error.empty() && a.GetVersion) a.GetVersion(&v);
  return v;
}

// Broadcast `bytes` of device memory from `root` on the context's stream (the fresh witness after the root's H2D copy).
int vimz_comm_broadcast_dev(vimz_ctx* ctx, vimz_comm* c, void* d_buf, size_t bytes, int root) {
  if (!ctx || !c || (!d_buf && bytes)) return set_error(VIMZ_ERR_ARG, "vimz_comm_broadcast_dev: null argument");
  if (ctx->device != c->device) return set_error(VIMZ_ERR_ARG, "vimz_comm_broadcast_dev: context and communicator on different devices");
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  DevGuard dg(ctx->device);
  if (c->world > 1 && bytes) VIMZ_NCCL(nccl().Broadcast(d_buf, d_buf, bytes, NCCL_UINT8, root, c->comm, ctx->stream));
  return VIMZ_OK;
}

// One MSM sharded by point range: this rank commits scalars[0 .. n) to ck[first .. first + n), the 96-byte partial sums are
// all-gathered on the context's stream and added on the GPU; every rank returns the full commitment.  One host wait.
int vimz_msm_sharded_dev(vimz_ctx* ctx, vimz_comm* c, const vimz_ck* ck, size_t first, const void* d_scalars, size_t n, vimz_point* out) {
  if (!ctx || !c || !ck || !out) return set_error(VIMZ_ERR_ARG, "vimz_msm_sharded_dev: null argument");
  if (ctx->device != c->device) return set_error(VIMZ_ERR_ARG, "vimz_msm_sharded_dev: context and communicator on different devices");
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  DevGuard dg(ctx->device);
  std::lock_guard<std::mutex> cl(c->mu);
  VIMZ_TRY(ctx->ws.result.reserve(4096));
  char* res = ctx->ws.result.as<char>();
  VIMZ_TRY(vimz_msm_async_dev(ctx, ck, first, d_scalars, n, res));
  const void* final_pt = res;
  if (c->world > 1) {
    VIMZ_NCCL(nccl().AllGather(res, c->d_gather, 96, NCCL_UINT8, c->comm, ctx->stream));
    VIMZ_TRY(curve_vtable(ctx->curve)->point_sum_batch(ctx, c->d_gather, (size_t)c->world, 1, res + 96));
    final_pt = res + 96;
  }
  VIMZ_CUDA(cudaMemcpyAsync(ctx->pinned, final_pt, 96, cudaMemcpyDeviceToHost, ctx->stream));
  VIMZ_CUDA(cudaStreamSynchronize(ctx->stream));
  memcpy(out, ctx->pinned, 96);
  return VIMZ_OK;
}

// One fold step of a row-range shard (vimz_acc_init_sharded) with the exchange inside: the step is enqueued, the rank's
// partial (comm_W2, comm_T) pair is all-gathered in stream order, the `world` pairs are added on the GPU and the host waits
// once for the two FULL commitments.  W2: device pointer valid on every rank (already broadcast / replicated).
int vimz_acc_step_begin_sharded_dev(vimz_acc* acc, vimz_comm* c, const void* d_W2, const vimz_fr* X2, vimz_point* comm_W2, vimz_point* comm_T) {
  if (!acc || !c || !comm_W2 || !comm_T) return set_error(VIMZ_ERR_ARG, "vimz_acc_step_begin_sharded_dev: null argument");
  vimz_ctx* ctx = acc->ctx;
  if (ctx->device != c->device) return set_error(VIMZ_ERR_ARG, "vimz_acc_step_begin_sharded_dev: accumulator and communicator on different devices");
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  DevGuard dg(ctx->device);
  std::lock_guard<std::mutex> cl(c->mu);
  void* d_part = nullptr;
  VIMZ_TRY(vimz_acc_step_begin_dev_async(acc, d_W2, X2, &d_part));
  if (c->world > 1) VIMZ_NCCL(nccl().AllGather(d_part, c->d_gather, 2 * 96, NCCL_UINT8, c->comm, ctx->stream));
  else VIMZ_CUDA(cudaMemcpyAsync(c->d_gather, d_part, 2 * 96, cudaMemcpyDeviceToDevice, ctx->stream));
  return vimz_acc_step_combine_dev(acc, c->d_gather, (size_t)c->world, comm_W2, comm_T);
}

// Same with the fresh witness in HOST memory on rank `root` only (W2 may be NULL elsewhere): the root copies it to its GPU,
// NCCL broadcasts it over NVLink into every rank's accumulator, then the step runs as above -- all in stream order, one wait.
int vimz_acc_step_begin_sharded(vimz_acc* acc, vimz_comm* c, const vimz_fr* W2, int root, const vimz_fr* X2, vimz_point* comm_W2,
                                vimz_point* comm_T) {
  if (!acc || !c || !comm_W2 || !comm_T) return set_error(VIMZ_ERR_ARG, "vimz_acc_step_begin_sharded: null argument");
  if (root < 0 || root >= c->world) return set_error(VIMZ_ERR_ARG, "vimz_acc_step_begin_sharded: bad root");
  if (c->rank == root && !W2 && acc->shape->n) return set_error(VIMZ_ERR_ARG, "vimz_acc_step_begin_sharded: the root rank must pass W2");
  vimz_ctx* ctx = acc->ctx;
  if (ctx->device != c->device) return set_error(VIMZ_ERR_ARG, "vimz_acc_step_begin_sharded: accumulator and communicator on different devices");
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  DevGuard dg(ctx->device);
  const size_t bytes = acc->shape->n * 32;
  {
    std::lock_guard<std::mutex> cl(c->mu);
    if (c->rank == root && bytes) VIMZ_CUDA(cudaMemcpyAsync(acc->W2, W2, bytes, cudaMemcpyHostToDevice, ctx->stream));
    if (c->world > 1 && bytes) VIMZ_NCCL(nccl().Broadcast(acc->W2, acc->W2, bytes, NCCL_UINT8, root, c->comm, ctx->stream));
  }
  return vimz_acc_step_begin_sharded_dev(acc, c, acc->W2, X2, comm_W2, comm_T);
}

}  // extern "C"
